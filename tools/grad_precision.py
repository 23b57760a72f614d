"""Backward-precision study: parameter gradients of every network of the compression model on the
B200 kernels against CPU autograd of the oracle, twice:

  fp32     the plain fp32 oracle (the reference's arithmetic) -- contains the ReLU-mask flips caused by the fp16-operand
           FORWARD of the product (activations that differ by 1e-3 around zero)
  matched  the oracle with its conv operands rounded to fp16 in the forward (rnd=round_fp16): the same arithmetic as the
           product's forward, but accumulation-order differences (1e-6) still move fp16 roundings of later layers, so its
           activations drift ~3-6e-4 away from the product's and a fraction of ReLU masks still differs
  forced   Encoder / Generator only: the matched oracle "teacher-forced" onto the product's own forward state -- every
           pre-norm conv output z of the oracle is replaced IN VALUE by the product's saved z (z + (z_product - z).detach(),
           so d z / d weights stays), hence every ChannelNorm statistic, ReLU mask and conv input of the oracle's backward
           is the product's.  What remains is purely the arithmetic of the product's BACKWARD kernels: the gradient
           operand format (fp16 vs bf16), fp32 accumulation order, split-K atomics

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

os.environ.setdefault("HFC_LPIPS_SYNTHETIC", "1")   # no checkpoints on the boxes: seeded stand-in, as the tests do

from hific_b200 import synth
from hific_b200.config import mse_lpips_args
from hific_b200.model import Model
from oracle import hific_oracle as O
import torch.nn.functional as F

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


def rows_to_nchw(rows, n, h, w, c):
    return rows.view(n, h, w, -1)[..., :c].permute(0, 3, 1, 2).float().cpu()


def forced_encoder(sd, x, zs, rnd):
    """O.encoder_forward (src/network/encoder.py:56-111) teacher-forced onto the product's pre-norm conv outputs."""
    p = "Encoder."
    force = lambda z, k: z + (zs[k] - z).detach()
    h = F.relu(O._cn(sd, p + "conv_block1.2", force(O._conv(sd, p + "conv_block1.1", O._reflect(x, (3, 3, 3, 3)), rnd=rnd), 0)))
    for i in range(2, 6):
        h = O._reflect(h, (0, 1, 1, 0))
        h = F.relu(O._cn(sd, p + f"conv_block{i}.2", force(O._conv(sd, p + f"conv_block{i}.1", h, stride=2, rnd=rnd), i - 1)))
    return O._conv(sd, p + "conv_block_out.1", O._reflect(h, (1, 1, 1, 1)), rnd=rnd)


def forced_generator(sd, y_hat, n_res, zs, rnd):
    """O.generator_forward (src/network/generator.py:98-169) teacher-forced onto the product's pre-norm conv outputs."""
    p = "Generator."
    force = lambda z, k: z + (zs[k] - z).detach()
    b1 = (1, 1, 1, 1)
    head = O._cn(sd, p + "conv_block_init.0", y_hat)
    head = O._cn(sd, p + "conv_block_init.3", force(O._conv(sd, p + "conv_block_init.2", O._reflect(head, b1), rnd=rnd), "init"))
    x = head
    for m in range(n_res):
        q = p + f"resblock_{m}"
        r = F.relu(O._cn(sd, q + ".norm1", force(O._conv(sd, q + ".conv1", O._reflect(x, b1), rnd=rnd), ("r", m, 0))))
        x = O._cn(sd, q + ".norm2", force(O._conv(sd, q + ".conv2", O._reflect(r, b1), rnd=rnd), ("r", m, 1))) + x
    x = x + head
    for i in range(1, 5):
        x = F.relu(O._cn(sd, p + f"upconv_block{i}.1", force(O._convT(sd, p + f"upconv_block{i}.0", x, 2, 1, 1, rnd=rnd), ("u", i))))
    return O._conv(sd, p + "conv_block_out.1", O._reflect(x, (3, 3, 3, 3)), rnd=rnd)


def study(name, module, prefix, sd, run_product, fn, inputs, up, forced=None):
    """run_product() -> (output, input grad[, saved state]) with .grad filled on the module's parameters."""
    for p in module.parameters():
        p.grad = None
    res = run_product()
    out_p, gin_p = res[0], res[1]
    row = {"network": name}
    variants = [("fp32", fn, O._ident), ("matched", fn, O.round_fp16)]
    if forced is not None:
        zs = res[2]
        variants.append(("forced", lambda s, t, r: forced(s, t, zs, r), O.round_fp16))
    for tag, fn, rnd in variants:
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
        xc = x.cuda()
        y = m.Encoder(xc)
        plan = m.Encoder._train_plans.get(xc)
        zs = {i: rows_to_nchw(z, 2, lay.oh, lay.ow, lay.cout) for i, (z, lay) in enumerate(zip(plan.z, plan.layers))}
        (y * up.cuda()).sum().backward()
        return y.detach(), None, zs
    rows.append(study("Encoder", m.Encoder, "Encoder.", sd, enc, lambda s, xx, r: O.encoder_forward(s, xx, rnd=r), [x], up,
                      forced=forced_encoder))

    yh = torch.round(2 * torch.randn(2, 220, 8, 8, generator=g))
    upx = torch.randn(2, 3, 128, 128, generator=g)

    def gen():
        yc = yh.cuda().requires_grad_(True)
        xh = m.Generator(yc)
        plan = m.Generator._train_plans.get(yc)
        zs = {"init": rows_to_nchw(plan.z_init, 2, 8, 8, 960)}
        for k, (z1, z2) in enumerate(plan.zr):
            zs[("r", k, 0)], zs[("r", k, 1)] = rows_to_nchw(z1, 2, 8, 8, 960), rows_to_nchw(z2, 2, 8, 8, 960)
        for i, (z, lay) in enumerate(zip(plan.zu, plan.ups)):
            zs[("u", i + 1)] = rows_to_nchw(z, 2, lay.oh, lay.ow, lay.cout)
        (xh * upx.cuda()).sum().backward()
        return xh.detach(), yc.grad, zs
    rows.append(study("Generator", m.Generator, "Generator.", sd, gen,
                      lambda s, t, r: O.generator_forward(s, t, n_residual_blocks=n_res, rnd=r), [yh], upx,
                      forced=lambda s, t, zs, r: forced_generator(s, t, n_res, zs, r)))

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
        for tag in ("fp32", "matched", "forced"):
            if tag not in r:
                continue
            e = r[tag]
            gi = f"{e['input_grad_rel_l2']:.2e}" if e["input_grad_rel_l2"] is not None else "-"
            print(f"{r['network']:16s} {tag:8s} {e['forward_rel_l2']:9.2e} {e['param_grads_rel_l2']:9.2e} "
                  f"{e['worst_tensor_rel_l2']:9.2e} {gi:>9s}  {e['worst_tensor']}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"grad_precision_{fmt}.json"), "w") as f:
        json.dump({"grad_fmt": fmt, "n_residual_blocks": n_res, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
