"""Kernel-level breakdown of one training step (CUPTI via torch.profiler: sees the libhfc launches as well as the
torch/cuDNN ones). Writes gpurun_out/train_profile.txt: per-kernel total time, launch count, share of the step.

    python tools/train_profile.py [--batch 32] [--gan]
"""
import argparse
import collections
import os
import sys

import torch

os.environ.setdefault("HFC_LPIPS_SYNTHETIC", "1")   # no checkpoints on the boxes: seeded stand-in, as the tests do

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hific_b200  # noqa: E402,F401
from hific_b200 import synth  # noqa: E402
from hific_b200.config import ModelModes, ModelTypes, mse_lpips_args, hific_args  # noqa: E402
from hific_b200.model import Model  # noqa: E402
from hific_b200.optim import Adam  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--gan", action="store_true")
    ap.add_argument("--out", default="gpurun_out/train_profile.txt")
    ap.add_argument("--detail", default="", help="comma-separated kernel-name substrings: list every launch (in order) with its duration")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = hific_args() if args.gan else mse_lpips_args()
    cfg.batch_size = args.batch
    cfg.image_dims = (3, 256, 256)
    import logging
    model = Model(cfg, logging.getLogger("tp"), model_mode=ModelModes.TRAINING,
                  model_type=ModelTypes.COMPRESSION_GAN if args.gan else ModelTypes.COMPRESSION)
    model.load_state_dict(synth.synth_state_dict(0), strict=False)
    model.to(dev).train()
    amort = [p for m in model.amortization_models for p in m.parameters()]
    hyper = list(model.Hyperprior.hyperlatent_likelihood.parameters())
    opt_a = Adam(amort, lr=1e-4)
    opt_h = Adam(hyper, lr=1e-4)
    x = synth.synth_image(args.batch, 256, 256, 1).to(dev)

    def step():
        losses = model(x, train_generator=True)
        losses['compression'].backward()
        opt_a.step(); opt_a.zero_grad()
        opt_h.step(); opt_h.zero_grad()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 3

    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = ev.name[:110]
            t, n = agg.get(k, (0.0, 0))
            agg[k] = (t + ev.device_time_total if hasattr(ev, "device_time_total") else t + ev.cuda_time_total, n + 1)
    tot = sum(t for t, _ in agg.values())
    detail = []
    if args.detail:
        pats = args.detail.split(",")
        evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and any(q in ev.name for q in pats)]
        evs.sort(key=lambda ev: ev.time_range.start)
        for ev in evs:
            detail.append(f"    {ev.device_time_total:9.1f} us  {ev.name[:60]}")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(f"# step {step_ms:.2f} ms wall (CUDA events, 3 steps); sum of kernel times {tot/1e3:.2f} ms; batch {args.batch}\n")
        for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{t/1e3:9.3f} ms {n:5d}x {100*t/tot:5.1f}%  {k}\n")
        if detail:
            f.write("# per-launch detail (launch order)\n" + "\n".join(detail) + "\n")
    print(open(args.out).read()[:6000])


if __name__ == "__main__":
    main()
