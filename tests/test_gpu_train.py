"""GPU tests of the training path (backward kernels) against torch autograd run on the CPU oracle.

Gradients are computed with bf16 GEMM operands (8-bit mantissa) on top of the fp16-operand forward, so the
tolerance is GRAD_TOL relative L2 per tensor -- an order of magnitude looser than the forward tolerance and stated
here because north_star only fixes the forward tolerance."""
import logging
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from hific_b200 import ops, synth  # noqa: E402
from hific_b200.config import mse_lpips_args  # noqa: E402
from hific_b200.model import Model  # noqa: E402
from oracle import hific_oracle as O  # noqa: E402
from test_gpu_parity import Feed, rel_l2  # noqa: E402

GRAD_TOL = 5e-2
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "grad_errors.txt")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def sd():
    return synth.synth_state_dict(0)


@pytest.fixture()
def model(sd):
    m = Model(mse_lpips_args(), logging.getLogger("train"))
    m.load_state_dict(sd, strict=True)
    return m.cuda().train()


def oracle_grads(sd, fn, inputs):
    """Run `fn(sd_with_grad, *inputs)` on CPU with autograd; returns (output, dict of parameter grads, input grads)."""
    torch.set_num_threads(os.cpu_count())
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ins = [t.clone().requires_grad_(True) for t in inputs]
    out = fn(sdg, *ins)
    return out, sdg, ins


def check_param_grads(module, prefix, sdg, tol=GRAD_TOL, skip=()):
    """Relative L2 error of every parameter gradient; all of them are logged before the worst one is asserted."""
    errs = []
    for name, p in module.named_parameters():
        key = prefix + name
        ref = sdg[key].grad
        if ref is None or any(s in key for s in skip):
            continue
        assert p.grad is not None, f"no gradient for {key}"
        errs.append((rel(p.grad, ref), key))
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            for e, k in errs:
                f.write(f"{e:.3e} {k}\n")
    except OSError:
        pass
    worst = max(errs)
    assert worst[0] < tol, f"{worst[1]}: rel err {worst[0]:.3e} (tolerance {tol})"
    return worst[0]


@pytest.mark.parametrize("c,n,h,w", [(960, 2, 8, 8), (480, 2, 9, 7), (240, 1, 12, 10), (220, 2, 8, 8), (120, 2, 13, 11),
                                     (60, 2, 17, 19), (12, 1, 9, 21)])
def test_channelnorm_backward_kernel(c, n, h, w):
    """Every (VEC, GROUP) instantiation of the backward kernel, odd pixel counts (partially filled warps), the fused
    bias gradient (column sums of dz), and the forward kernel's fp32 output at the same widths."""
    g = torch.Generator().manual_seed(1)
    z = (torch.randn(n, c, h, w, generator=g) * 2 + 0.3).cuda().requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(1, c, 1, 1, generator=g)).cuda().requires_grad_(True)
    beta = (0.1 * torch.randn(1, c, 1, 1, generator=g)).cuda().requires_grad_(True)
    up = torch.randn(n, c, h, w, generator=g).cuda()
    from hific_b200.train_plan import nchw_to_rows, norm_bwd, norm_fwd
    for act in (ops.ACT_NONE, ops.ACT_RELU):
        for t in (z, gamma, beta):
            t.grad = None
        y = O.channel_norm(z, gamma, beta)
        if act == ops.ACT_RELU:
            y = torch.relu(y)
        y.backward(up)
        dz, dg, db, dbias = norm_bwd(nchw_to_rows(z.detach()), nchw_to_rows(up), gamma.detach(), beta.detach(), act,
                                     as_operand=False)
        assert rel(dz.view(n, h, w, c).permute(0, 3, 1, 2), z.grad) < 1e-4
        # the same result written directly as the bf16 GEMM operand (pitch round_up(c, 64), zero channel padding)
        op, dg2, _, dbias2 = norm_bwd(nchw_to_rows(z.detach()), nchw_to_rows(up), gamma.detach(), beta.detach(), act)
        from hific_b200.grad import GRAD_BF16
        fmt = torch.bfloat16 if GRAD_BF16 else torch.float16
        op = op.clone().view(fmt)
        assert op.shape == (n * h * w, ops.round_up(c, 64))
        assert torch.equal(op[:, :c], dz[:, :c].to(fmt))
        assert float(op[:, c:].float().abs().max() if op.shape[1] > c else 0.0) == 0.0
        assert rel(dg2, dg) < 1e-6 and rel(dbias2, dbias) < 1e-5
        assert rel(dg, gamma.grad) < 1e-4 and rel(db, beta.grad) < 1e-4
        assert rel(dbias, z.grad.sum(dim=(0, 2, 3))) < 1e-3 or float(z.grad.sum(dim=(0, 2, 3)).norm()) < 1e-3
        geom = ops.Geom(n, h, w, c, ops.round_up(c, 64), 1, 1, 1, 1)
        act_buf, f32 = norm_fwd(nchw_to_rows(z.detach()), geom, gamma.detach(), beta.detach(), act, True, want_f32=True)
        assert rel(f32.view(n, h, w, c).permute(0, 3, 1, 2), y) < 1e-5
        inner = act_buf[:, 1:-1, 1:-1, :c].float().permute(0, 3, 1, 2)
        assert rel(inner, y) < 1e-3
        assert float(act_buf[..., c:].abs().max() if geom.cpad > c else 0.0) == 0.0
        assert torch.equal(act_buf[:, 0, 1:-1], act_buf[:, 2, 1:-1])            # reflected top row


def test_likelihood_backward_kernels(sd):
    g = torch.Generator().manual_seed(2)
    shape = (2, 220, 8, 8)
    y = (2 * torch.randn(shape, generator=g)).cuda().requires_grad_(True)
    mu = torch.randn(shape, generator=g).cuda().requires_grad_(True)
    sr = (2 * torch.rand(shape, generator=g)).cuda().requires_grad_(True)
    nz = (torch.rand(shape, generator=g) - 0.5).cuda()
    up = torch.randn(shape, generator=g).cuda()
    dec, sums = ops.LatentLikelihoodFn.apply(y, mu, sr, nz, 0.11, "gaussian")
    (sums[0].float() * 0.37 + (dec * up).sum()).backward()
    got = [t.grad.clone() for t in (y, mu, sr)]
    for t in (y, mu, sr):
        t.grad = None
    from hific_b200.hyperprior import MIN_SCALE
    import sys
    sys.path.insert(0, "/nonexistent")
    # torch reference with the reference's LowerBoundToward semantics
    class LBT(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t, b):
            ctx.mask = t.ge(b)
            return torch.clamp(t, b)

        @staticmethod
        def backward(ctx, go):
            return go * torch.logical_or(ctx.mask, go.lt(0.)).type(go.dtype), None
    sc = LBT.apply(sr, 0.11)
    d = (y + nz - mu).abs()
    cdf = lambda v: 0.5 * torch.erfc(v * (-1.0 / 2 ** 0.5))
    p = LBT.apply(cdf((0.5 - d) / sc) - cdf(-(0.5 + d) / sc), 1e-9)
    v = y - mu
    dec_ref = v + (torch.floor(v + 0.5) - v).detach() + mu
    (torch.log(p + 1e-9).sum() * 0.37 + (dec_ref * up).sum()).backward()
    for a, t, name in zip(got, (y, mu, sr), ("dy", "dmu", "dscale")):
        assert rel(a, t.grad) < 1e-4, name


def test_density_backward_kernel(sd):
    g = torch.Generator().manual_seed(3)
    C = 320
    keys = [f"Hyperprior.hyperlatent_likelihood.{n}_{k}" for k in range(4) for n in ("H", "a", "b")]
    Hs = [sd[f"Hyperprior.hyperlatent_likelihood.H_{k}"].clone().cuda().requires_grad_(True) for k in range(4)]
    a_s = [sd[f"Hyperprior.hyperlatent_likelihood.a_{k}"].clone().cuda().requires_grad_(True) for k in range(4)]
    bs = [sd[f"Hyperprior.hyperlatent_likelihood.b_{k}"].clone().cuda().requires_grad_(True) for k in range(4)]
    z = (2 * torch.randn(3, C, 4, 4, generator=g)).cuda().requires_grad_(True)
    nz = (torch.rand(3, C, 4, 4, generator=g) - 0.5).cuda()
    up = torch.randn(3, C, 4, 4, generator=g).cuda()
    packed = ops.pack_density_params_autograd(Hs, a_s, bs)
    zn, zq, sums = ops.HyperlatentLikelihoodFn.apply(z, packed, nz)
    (sums[0].float() * 0.21 + (zn * up).sum()).backward()
    got = [t.grad.clone() for t in [z] + Hs + a_s + bs]
    for t in [z] + Hs + a_s + bs:
        t.grad = None
    sdl = {}
    for k in range(4):
        sdl[f"H_{k}"], sdl[f"a_{k}"], sdl[f"b_{k}"] = Hs[k], a_s[k], bs[k]
    lik = O.density_likelihood(sdl, z + nz, prefix="")
    # O.density_likelihood clamps without the LowerBoundToward gradient gate; no element is at the floor here
    assert (lik > 2e-9).all()
    (torch.log(lik + 1e-9).sum() * 0.21 + ((z + nz) * up).sum()).backward()
    for a, t in zip(got, [z] + Hs + a_s + bs):
        assert rel(a, t.grad) < 2e-4


def test_encoder_backward_vs_oracle(model, sd):
    x = synth.synth_image(2, 64, 64, 0)
    g = torch.Generator().manual_seed(4)
    up = torch.randn(2, 220, 4, 4, generator=g)
    y = model.Encoder(x.cuda())
    assert y.requires_grad
    (y * up.cuda()).sum().backward()
    out, sdg, _ = oracle_grads(sd, lambda s, xx: O.encoder_forward(s, xx), [x])
    (out * up).sum().backward()
    assert rel(y, out) < 1e-3
    check_param_grads(model.Encoder, "Encoder.", sdg)


def test_hyper_networks_backward_vs_oracle(model, sd):
    g = torch.Generator().manual_seed(5)
    y = torch.randn(2, 220, 16, 16, generator=g)
    upz = torch.randn(2, 320, 4, 4, generator=g)
    yc = y.cuda().requires_grad_(True)
    z = model.Hyperprior.analysis_net(yc)
    (z * upz.cuda()).sum().backward()
    out, sdg, ins = oracle_grads(sd, lambda s, yy: O.hyper_analysis(s, yy), [y])
    (out * upz).sum().backward()
    check_param_grads(model.Hyperprior.analysis_net, "Hyperprior.analysis_net.", sdg)
    assert rel(yc.grad, ins[0].grad) < GRAD_TOL
    zz = torch.randn(2, 320, 4, 4, generator=g)
    upm = torch.randn(2, 220, 16, 16, generator=g)
    zc = zz.cuda().requires_grad_(True)
    mu = model.Hyperprior.synthesis_mu(zc)
    (mu * upm.cuda()).sum().backward()
    out, sdg, ins = oracle_grads(sd, lambda s, t: O.hyper_synthesis(s, t, "Hyperprior.synthesis_mu."), [zz])
    (out * upm).sum().backward()
    check_param_grads(model.Hyperprior.synthesis_mu, "Hyperprior.synthesis_mu.", sdg)
    assert rel(zc.grad, ins[0].grad) < GRAD_TOL


def test_generator_backward_vs_oracle(sd):
    cfg = mse_lpips_args()
    cfg.n_residual_blocks = 2                                   # keeps the CPU autograd reference quick
    sd2 = {k: v for k, v in synth.synth_state_dict(0, n_residual_blocks=2).items()}
    m = Model(cfg, logging.getLogger("g2"))
    m.load_state_dict(sd2, strict=True)
    m.cuda().train()
    g = torch.Generator().manual_seed(6)
    yh = torch.round(2 * torch.randn(1, 220, 4, 4, generator=g))
    up = torch.randn(1, 3, 64, 64, generator=g)
    yc = yh.cuda().requires_grad_(True)
    xh = m.Generator(yc)
    (xh * up.cuda()).sum().backward()
    out, sdg, ins = oracle_grads(sd2, lambda s, t: O.generator_forward(s, t, n_residual_blocks=2), [yh])
    (out * up).sum().backward()
    assert rel(xh, out) < 2e-3
    check_param_grads(m.Generator, "Generator.", sdg)
    assert rel(yc.grad, ins[0].grad) < GRAD_TOL


def test_full_training_step_vs_oracle(sd):
    """loss = rate + k_M * distortion (LPIPS excluded here: its trunk runs on cuDNN): every parameter gradient of the
    compression model against CPU autograd of the oracle, same noise."""
    cfg = mse_lpips_args()
    cfg.n_residual_blocks = 2
    sd2 = synth.synth_state_dict(0, n_residual_blocks=2)
    m = Model(cfg, logging.getLogger("full"))
    m.load_state_dict(sd2, strict=True)
    m.cuda().train()
    x = synth.synth_image(2, 128, 128, 0)
    nz = synth.synth_noise((2, 320, 2, 2), "zt", 0)
    ny = synth.synth_noise((2, 220, 8, 8), "yt", 0)
    with Feed([nz, ny]):
        inter, info = m.compression_forward(x.cuda())
    loss = 2.0 * inter.n_bpp + cfg.k_M * m.distortion_loss(inter.reconstruction, inter.input_image)
    loss.backward()

    def fwd(s, xx):
        recon, hyper, _ = O.compression_forward(s, xx, True, False, nz, ny, n_residual_blocks=2)
        return 2.0 * hyper.total_nbpp + cfg.k_M * O.distortion_loss(recon, xx)
    out, sdg, _ = oracle_grads(sd2, fwd, [x])
    out.backward()
    assert abs(float(loss) - float(out)) < 0.05 * abs(float(out))
    for name, prm in m.Hyperprior.named_parameters():
        ref = sdg["Hyperprior." + name].grad
        print(f"{name:40s} |ours| {prm.grad.norm().item():.4e} |ref| {ref.norm().item():.4e} rel {rel(prm.grad, ref):.3e}")
    # y_hat rounding flips perturb the generator-side gradients (the loss surface is only piecewise smooth): the
    # hyperprior / rate side is compared strictly, the rest loosely
    check_param_grads(m.Hyperprior, "Hyperprior.", sdg, tol=0.1)
    check_param_grads(m.Encoder, "Encoder.", sdg, tol=0.3)
    check_param_grads(m.Generator, "Generator.", sdg, tol=0.3)


def test_backward_arithmetic_at_the_products_own_forward_state(sd):
    """The backward kernels against fp32 autograd evaluated AT THE SAME forward state (tools/grad_precision.py, "forced"):
    the oracle's pre-norm conv outputs are replaced in value by the product's saved ones, so ChannelNorm statistics, ReLU
    masks and conv inputs of the two backward passes coincide and no mask flip separates them (the network-level bars
    above are dominated by flips that ANY 10-bit forward has against an fp32 one: cuDNN's TF32 path shows 1.5-4.8e-2 on
    the LPIPS trunk, gpurun_out/lpips_trunk_errors.txt).  What is left is the arithmetic of the backward GEMMs:
    fp16 gradient operands (default) must hold every parameter gradient of Encoder and Generator to 2e-3 relative L2,
    the bf16 operands of round 1 to 1.5e-2 (measured through the kernel emulation: 8e-4 / 7e-3)."""
    from tools import grad_precision as GP
    from hific_b200.grad import GRAD_BF16
    tol = 1.5e-2 if GRAD_BF16 else 2e-3
    n_res = 2
    cfg = mse_lpips_args()
    cfg.n_residual_blocks = n_res
    sd2 = synth.synth_state_dict(0, n_residual_blocks=n_res)
    m = Model(cfg, logging.getLogger("forced"))
    m.load_state_dict(sd2, strict=True)
    m.cuda().train()
    g = torch.Generator().manual_seed(11)
    x = synth.synth_image(2, 128, 128, 0)
    up = torch.randn(2, 220, 8, 8, generator=g)
    yh = torch.round(2 * torch.randn(2, 220, 8, 8, generator=g))
    upx = torch.randn(2, 3, 128, 128, generator=g)

    def enc():
        xc = x.cuda()
        y = m.Encoder(xc)
        plan = m.Encoder._train_plans.get(xc)
        zs = {i: GP.rows_to_nchw(z, 2, lay.oh, lay.ow, lay.cout) for i, (z, lay) in enumerate(zip(plan.z, plan.layers))}
        (y * up.cuda()).sum().backward()
        return y.detach(), None, zs

    def gen():
        yc = yh.cuda().requires_grad_(True)
        xh = m.Generator(yc)
        plan = m.Generator._train_plans.get(yc)
        zs = {"init": GP.rows_to_nchw(plan.z_init, 2, 8, 8, 960)}
        for k, (z1, z2) in enumerate(plan.zr):
            zs[("r", k, 0)], zs[("r", k, 1)] = GP.rows_to_nchw(z1, 2, 8, 8, 960), GP.rows_to_nchw(z2, 2, 8, 8, 960)
        for i, (z, lay) in enumerate(zip(plan.zu, plan.ups)):
            zs[("u", i + 1)] = GP.rows_to_nchw(z, 2, lay.oh, lay.ow, lay.cout)
        (xh * upx.cuda()).sum().backward()
        return xh.detach(), yc.grad, zs
    re = GP.study("Encoder", m.Encoder, "Encoder.", sd2, enc, lambda s, t, r: O.encoder_forward(s, t, rnd=r), [x], up,
                  forced=GP.forced_encoder)
    rg = GP.study("Generator", m.Generator, "Generator.", sd2, gen,
                  lambda s, t, r: O.generator_forward(s, t, n_residual_blocks=n_res, rnd=r), [yh], upx,
                  forced=lambda s, t, zs, r: GP.forced_generator(s, t, n_res, zs, r))
    for r in (re, rg):
        f = r["forced"]
        print(r["network"], {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in f.items()})
        assert f["forward_rel_l2"] < 1e-4                   # the oracle really sits on the product's forward state
        assert f["worst_tensor_rel_l2"] < tol, (r["network"], f)
        if f["input_grad_rel_l2"] is not None:
            assert f["input_grad_rel_l2"] < tol
