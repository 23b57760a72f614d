"""Stage-wise parity of the CUDA path against the CPU oracle (run on the GPU box).

Protocol ('quantisation discontinuity'): each stage is fed the ORACLE's input for that
stage, so a rounding flip in y_hat does not masquerade as a generator error; the end-to-end numbers are
reported separately.  Prints rel-L2 / max-abs per stage and writes gpurun_out/parity_report.json.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import logging

import torch

os.environ.setdefault("HFC_LPIPS_SYNTHETIC", "1")   # no checkpoints on the boxes: seeded stand-in, as the tests do

from hific_b200 import synth
from hific_b200.config import ModelModes, ModelTypes, mse_lpips_args
from hific_b200.model import Model
from oracle import hific_oracle as O


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def max_abs(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item()


class Feed:
    def __init__(self, noises):
        self.noises, self.calls = list(noises), 0

    def __enter__(self):
        self._orig = torch.nn.init.uniform_

        def fake(t, a=0.0, b=1.0):
            if (a, b) != (-0.5, 0.5) or self.calls >= len(self.noises):
                return self._orig(t, a, b)
            n = self.noises[self.calls]
            self.calls += 1
            with torch.no_grad():
                t.copy_(n.to(t.device))
            return t

        torch.nn.init.uniform_ = fake
        return self

    def __exit__(self, *e):
        torch.nn.init.uniform_ = self._orig


def stage_report(model, sd, b, h, w, training, seed=0, rnd=None):
    x = synth.synth_image(b, h, w, seed)
    noise_z = synth.synth_noise((b, 320, h // 64, w // 64), "pz", seed)
    noise_y = synth.synth_noise((b, 220, h // 16, w // 16), "py", seed)
    with torch.no_grad():
        recon_o, hyp_o, y_o = O.compression_forward(sd, x, training, False, noise_z, noise_y)
    model.train(training)
    dev = "cuda"
    rep = {}
    with torch.no_grad():
        y = model.Encoder(x.to(dev))
        rep["encoder_y_rel_l2"] = rel_l2(y, y_o)
        rep["encoder_y_max_abs"] = max_abs(y, y_o)
        z = model.Hyperprior.analysis_net(y_o.to(dev))
        rep["analysis_z_rel_l2"] = rel_l2(z, hyp_o.hyperlatents)
        z_dec_o = hyp_o.noisy_hyperlatents if training else hyp_o.quantized_hyperlatents
        mu = model.Hyperprior.synthesis_mu(z_dec_o.to(dev))
        sg = model.Hyperprior.synthesis_std(z_dec_o.to(dev))
        rep["synthesis_mu_rel_l2"] = rel_l2(mu, hyp_o.latent_means)
        rep["synthesis_sigma_rel_l2"] = rel_l2(sg.clamp(min=0.11), hyp_o.latent_scales)
        # hyperprior as a whole, fed the oracle's y
        with Feed([noise_z, noise_y]):
            info = model.Hyperprior(y_o.to(dev), spatial_shape=(h, w))
        flips = ((info.decoded.cpu() - hyp_o.decoded).abs() > 0.5).float().mean().item()
        rep["yhat_mismatch_fraction"] = flips
        for f in ("latent_nbpp", "hyperlatent_nbpp", "latent_qbpp", "hyperlatent_qbpp", "total_nbpp", "total_qbpp"):
            rep[f + "_abs_err"] = abs(float(getattr(info, f)) - float(getattr(hyp_o, f)))
            rep[f + "_oracle"] = float(getattr(hyp_o, f))
        xh = model.Generator(hyp_o.decoded.to(dev))
        rep["generator_xhat_rel_l2"] = rel_l2(xh, recon_o)
        rep["generator_xhat_max_abs"] = max_abs(xh, recon_o)
        rep["xhat_oracle_absmax"] = recon_o.abs().max().item()
        # end to end
        with Feed([noise_z, noise_y]):
            inter, info2 = model.compression_forward(x.to(dev))
        rep["e2e_xhat_rel_l2"] = rel_l2(inter.reconstruction, recon_o)
        rep["e2e_total_qbpp_abs_err"] = abs(float(info2.total_qbpp) - float(hyp_o.total_qbpp))
        rep["e2e_total_nbpp_abs_err"] = abs(float(info2.total_nbpp) - float(hyp_o.total_nbpp))
    return rep


def main():
    torch.set_num_threads(os.cpu_count())
    sd = synth.synth_state_dict(0)
    model = Model(mse_lpips_args(), logging.getLogger("parity"))
    model.load_state_dict(sd, strict=True)
    model.cuda()
    out = {}
    for name, (b, h, w, training) in {"train_2x128": (2, 128, 128, True), "eval_1x256": (1, 256, 256, False)}.items():
        out[name] = stage_report(model, sd, b, h, w, training)
        print(name)
        for k, v in out[name].items():
            print(f"   {k:32s} {v:.4e}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/parity_report.json", "w"), indent=1)


if __name__ == "__main__":
    main()
