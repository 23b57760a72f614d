"""A/B timing of the three schedules of the Gaussian latent-likelihood kernel (HFC_LIKELIHOOD_V=1|2|3) at the c2
(1 802 240 elements) and c5 (7 208 960) sizes: CUDA events on the launching stream, L2 flushed before every launch,
20 B/element algorithmic traffic against the measured HBM peak.  GPU box only; prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hific_b200 import ops


def hbm_peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def main():
    peak = hbm_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"hbm_peak_gbs": peak, "bytes_per_element": 20, "sizes": {}}
    for name, n in (("c2", 1802240), ("c5", 7208960)):
        g = torch.Generator(device="cuda").manual_seed(0)
        y = torch.randn(n, device="cuda", generator=g).view(1, 1, 1, n) * 2
        mu = torch.randn(n, device="cuda", generator=g).view_as(y)
        s = torch.rand(n, device="cuda", generator=g).view_as(y) * 2
        nz = torch.rand(n, device="cuda", generator=g).view_as(y) - 0.5
        res = {}
        for v in ("1", "2", "3"):
            os.environ["HFC_LIKELIHOOD_V"] = v
            sums = torch.zeros(2, dtype=torch.float64, device="cuda")
            for _ in range(5):
                ops.latent_likelihood(y, mu, s, nz, sums=sums)
            cold, warm = [], []
            for flushed in (True, False):
                for _ in range(30):
                    if flushed:
                        flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.latent_likelihood(y, mu, s, nz, sums=sums)
                    e1.record()
                    torch.cuda.synchronize()
                    (cold if flushed else warm).append(e0.elapsed_time(e1) * 1e3)
            cold.sort(); warm.sort()
            us = cold[len(cold) // 2]
            res["v" + v] = {"us_cold_l2_median": us, "us_cold_l2_min": cold[0], "us_warm_median": warm[len(warm) // 2],
                            "gbs": 20.0 * n / (us * 1e-6) / 1e9, "frac_of_hbm_peak": 20.0 * n / (us * 1e-6) / 1e9 / peak}
        out["sizes"][name] = res
    os.environ.pop("HFC_LIKELIHOOD_V", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
