"""Backward-precision study: parameter gradients of every network of the compression model on the
B200 kernels against CPU autograd of the oracle, twice:

  fp32     the plain fp32 oracle (the reference's arithmetic) -- contains the ReLU-mask flips caused by the fp16-operand
           FORWARD of the product (activations that differ by 1e-3 around zero)
  matched  the oracle with its conv operands rounded to fp16 in the forward (rnd=round_fp16): same activations and masks
           as the product's forward up to accumulation order, fp32 backward arithmetic on the same saved operands --
           what remains is the arithmetic error of the product's BACKWARD GEMMs (gradient operand format)

The gradient operand format is selected with HFC_GRAD_FMT (bf16 | fp16); the script prints one table per run and writes
gpurun_out/grad_precision_<fmt>.json.  Full 9-residual-block architecture at 2 x 128 x 128.  GPU box only.
"""
import json
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from hific_b200 import synth
from hific_b200.config import mse_lpips_args
from hific_b200.model import Model
from oracle import hific_oracle as O

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.set_num_threads(os.cpu_count())


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def oracle_grads(sd, fn, inputs, rnd):
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ins = [t.clone().requires_grad_(True) for t in inputs]
    out = fn(sdg, *ins, rnd)
    return out, sdg, ins


def compare(module, prefix, sdg):
    num = den = 0.0
    worst = (0.0, "")
    for name, p in module.named_parameters():
        ref = sdg[prefix + name].grad
        if ref is None:
            continue
        d = (p.grad.double().cpu() - ref.double())
        num += float(d.pow(2).sum())
        den += float(ref.double().pow(2).sum())
        worst = max(worst, (float(d.norm() / ref.double().norm().clamp_min(1e-300)), name))
    return (num / den) ** 0.5, worst


def study(name, module, prefix, sd, run_product, fn, inputs, up):
    """run_product() -> (output, input grad) with .grad filled on the module's parameters."""
    for p in module.parameters():
        p.grad = None
    out_p, gin_p = run_product()
    row = {"network": name}
    for tag, rnd in (("fp32", O._ident), ("matched", O.round_fp16)):
        out, sdg, ins = oracle_grads(sd, fn, inputs, rnd)
        (out * up).sum().backward()
        agg, worst = compare(module, prefix, sdg)
        row[tag] = {"forward_rel_l2": rel(out_p, out.detach()), "param_grads_rel_l2": agg, "worst_tensor_rel_l2": worst[0],
                    "worst_tensor": worst[1], "input_grad_rel_l2": rel(gin_p, ins[0].grad) if gin_p is not None else None}
    return row


def main():
    fmt = os.environ.get("HFC_GRAD_FMT", "default")
    n_res = int(os.environ.get("HFC_NRES", 9))
    cfg = mse_lpips_args()
    cfg.n_residual_blocks = n_res
    sd = synth.synth_state_dict(0, n_residual_blocks=n_res)
    m = Model(cfg, logging.getLogger("prec"))
    m.load_state_dict(sd, strict=True)
    m.cuda().train()
    g = torch.Generator().manual_seed(11)
    rows = []

    x = synth.synth_image(2, 128, 128, 0)
    up = torch.randn(2, 220, 8, 8, generator=g)

    def enc():
        y = m.Encoder(x.cuda())
        (y * up.cuda()).sum().backward()
        return y.detach(), None
    rows.append(study("Encoder", m.Encoder, "Encoder.", sd, enc, lambda s, xx, r: O.encoder_forward(s, xx, rnd=r), [x], up))

    yh = torch.round(2 * torch.randn(2, 220, 8, 8, generator=g))
    upx = torch.randn(2, 3, 128, 128, generator=g)

    def gen():
        yc = yh.cuda().requires_grad_(True)
        xh = m.Generator(yc)
        (xh * upx.cuda()).sum().backward()
        return xh.detach(), yc.grad
    rows.append(study("Generator", m.Generator, "Generator.", sd, gen,
                      lambda s, t, r: O.generator_forward(s, t, n_residual_blocks=n_res, rnd=r), [yh], upx))

    y = torch.randn(2, 220, 16, 16, generator=g)
    upz = torch.randn(2, 320, 4, 4, generator=g)

    def ana():
        yc = y.cuda().requires_grad_(True)
        z = m.Hyperprior.analysis_net(yc)
        (z * upz.cuda()).sum().backward()
        return z.detach(), yc.grad
    rows.append(study("HyperAnalysis", m.Hyperprior.analysis_net, "Hyperprior.analysis_net.", sd, ana,
                      lambda s, t, r: O.hyper_analysis(s, t, rnd=r), [y], upz))

    zz = torch.randn(2, 320, 4, 4, generator=g)
    upm = torch.randn(2, 220, 16, 16, generator=g)

    def syn():
        zc = zz.cuda().requires_grad_(True)
        mu = m.Hyperprior.synthesis_mu(zc)
        (mu * upm.cuda()).sum().backward()
        return mu.detach(), zc.grad
    rows.append(study("HyperSynthesis", m.Hyperprior.synthesis_mu, "Hyperprior.synthesis_mu.", sd, syn,
                      lambda s, t, r: O.hyper_synthesis(s, t, "Hyperprior.synthesis_mu.", rnd=r), [zz], upm))

    print(f"gradient operand format: {fmt}   ({n_res} residual blocks, 2 x 128 x 128)")
    print(f"{'network':16s} {'oracle':8s} {'fwd':>9s} {'params':>9s} {'worst':>9s} {'d input':>9s}  worst tensor")
    for r in rows:
        for tag in ("fp32", "matched"):
            e = r[tag]
            gi = f"{e['input_grad_rel_l2']:.2e}" if e["input_grad_rel_l2"] is not None else "-"
            print(f"{r['network']:16s} {tag:8s} {e['forward_rel_l2']:9.2e} {e['param_grads_rel_l2']:9.2e} "
                  f"{e['worst_tensor_rel_l2']:9.2e} {gi:>9s}  {e['worst_tensor']}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"grad_precision_{fmt}.json"), "w") as f:
        json.dump({"grad_fmt": fmt, "n_residual_blocks": n_res, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
