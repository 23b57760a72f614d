"""CPU: the TRAINING plans of hific_b200.train_plan (what is saved in the forward pass, the order of the adjoints, which
gradient lands in which parameter slot, the residual / head fan-out of the Generator) against torch autograd of the
oracle, with the kernels replaced by torch stand-ins (tests/emulation.py).  Host logic only; the kernels are checked by
tests/test_gpu_grad.py / test_gpu_train.py on a GPU.  Tolerance: bf16 gradient operands, fp16 activations, ReLU flips of
those activations -> 5e-2 relative L2 per tensor (the bar of the GPU tests)."""
import pytest
import torch

from emulation import training_cpu_emulation
from hific_b200 import synth
from hific_b200.network import encoder, generator, hyper
from oracle import hific_oracle as O

TOL = 8e-2          # small maps: ReLU-flip noise averages over fewer pixels than in the GPU tests (5e-2 on 128 x 128 x 2)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def sd():
    return synth.synth_state_dict(0)


def check(module, prefix, sd, x, oracle_fn, out_weight, input_grad=True):
    module.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, strict=True)
    module.train()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(prefix)}
    xo = x.clone().requires_grad_(True)
    (oracle_fn(sdg, xo) * out_weight).sum().backward()
    xp = x.clone().requires_grad_(True)
    with training_cpu_emulation():
        (module(xp) * out_weight).sum().backward()
    if input_grad:
        assert rel(xp.grad, xo.grad) < TOL, "input gradient"
    else:
        assert xp.grad is None                   # the image needs no gradient: the plan skips the first data gradient
    worst = ("", 0.0)
    for name, p in module.named_parameters():
        r = rel(p.grad, sdg[prefix + name].grad)
        worst = max(worst, (name, r), key=lambda t: t[1])
    assert worst[1] < TOL, worst


def test_encoder_training_plan(sd):
    g = torch.Generator().manual_seed(0)
    x = synth.synth_image(1, 64, 64, 2)
    wgt = torch.randn((1, 220, 4, 4), generator=g)
    check(encoder.Encoder((3, 64, 64), 1, C=220), "Encoder.", sd, x, lambda s, t: O.encoder_forward(s, t), wgt,
          input_grad=False)


def test_generator_training_plan(sd):
    g = torch.Generator().manual_seed(1)
    y = torch.round(torch.randn((1, 220, 2, 3), generator=g) * 2)
    wgt = torch.randn((1, 3, 32, 48), generator=g)
    check(generator.Generator((220, 2, 3), 1, C=220, n_residual_blocks=9), "Generator.", sd, y,
          lambda s, t: O.generator_forward(s, t), wgt)


def test_hyper_training_plans(sd):
    g = torch.Generator().manual_seed(2)
    y = torch.randn((2, 220, 16, 16), generator=g)
    check(hyper.HyperpriorAnalysis(C=220, N=320), "Hyperprior.analysis_net.", sd, y,
          lambda s, t: O.hyper_analysis(s, t), torch.randn((2, 320, 4, 4), generator=g))
    z = torch.round(torch.randn((2, 320, 2, 3), generator=g) * 3)
    check(hyper.HyperpriorSynthesis(C=220, N=320), "Hyperprior.synthesis_std.", sd, z,
          lambda s, t: O.hyper_synthesis(s, t, "Hyperprior.synthesis_std."), torch.randn((2, 220, 8, 12), generator=g))


@pytest.mark.parametrize("channel_norm", [True, False], ids=["channel-norm", "instance-norm"])
def test_whole_training_step_on_cpu(channel_norm):
    """(instance-norm: the same step with `args.use_channel_norm = False`, src/model.py:69-72 -> InstanceNorm2d in both networks.)
    CPU mirror of tests/test_gpu_train.py::test_full_training_step_vs_oracle: `Model.compression_forward` in training
    mode (noise fed as the reference draws it), rate + distortion loss, `backward()` through every training plan and
    autograd Function -- every parameter gradient against autograd of the oracle.  Same tolerances as on the GPU: the rate
    side strictly, the image side loosely (rounding flips of y_hat make the loss only piecewise smooth)."""
    import logging
    from emulation import train_step_cpu_emulation
    from hific_b200.config import mse_lpips_args
    from hific_b200.model import Model
    from oracle.ref_shim import NoiseFeeder
    cfg = mse_lpips_args()
    cfg.n_residual_blocks = 2
    cfg.use_channel_norm = channel_norm
    sd2 = synth.synth_state_dict(0, n_residual_blocks=2)
    if not channel_norm:
        sd2 = synth.instance_norm_variant(sd2)
    m = Model(cfg, logging.getLogger("cpu-train"))
    m.load_state_dict(sd2, strict=True)
    m.train()
    x = synth.synth_image(2, 128, 128, 0)
    nz = synth.synth_noise((2, 320, 2, 2), "zt", 0)
    ny = synth.synth_noise((2, 220, 8, 8), "yt", 0)
    with train_step_cpu_emulation():
        with NoiseFeeder([nz, ny]):
            inter, info = m.compression_forward(x)
        loss = 2.0 * inter.n_bpp + cfg.k_M * m.distortion_loss(inter.reconstruction, inter.input_image)
        loss.backward()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd2.items()}
    recon, hyp, _ = O.compression_forward(sdg, x, True, False, nz, ny, n_residual_blocks=2)
    out = 2.0 * hyp.total_nbpp + cfg.k_M * O.distortion_loss(recon, x)
    out.backward()
    assert abs(float(loss) - float(out)) < 0.05 * abs(float(out))

    def worst(module, prefix):
        w = ("", 0.0)
        for name, p in module.named_parameters():
            assert p.grad is not None, prefix + name
            want = sdg[prefix + name].grad
            if not channel_norm and name.endswith(".bias") and prefix + name[:-5] + ".weight" in sdg and \
                    want.abs().max() < 1e-3 * sdg[prefix + name[:-5] + ".weight"].grad.abs().max():
                continue     # conv bias in front of an InstanceNorm: mathematically zero gradient, rounding noise on both sides
            w = max(w, (name, rel(p.grad, want)), key=lambda t: t[1])
        return w
    assert worst(m.Hyperprior, "Hyperprior.")[1] < 0.1
    # image side: with these synthetic weights the instance-norm network is far more sensitive to y_hat rounding flips than
    # the channel-norm one -- the ORACLE with fp16-rounded conv operands differs from its own fp32 run by 0.45 (Encoder,
    # conv_block1.2.weight) where the channel-norm oracle differs by 0.08; the product sits at 0.41 / 0.49 (measured)
    image_side = 0.3 if channel_norm else 0.6
    assert worst(m.Encoder, "Encoder.")[1] < image_side
    assert worst(m.Generator, "Generator.")[1] < image_side


def test_backward_arithmetic_at_the_products_own_forward_state_cpu():
    """CPU twin (kernel emulation) of tests/test_gpu_train.py::test_backward_arithmetic_at_the_products_own_forward_state:
    with the oracle teacher-forced onto the product's forward state the emulated fp16-operand backward agrees with fp32
    autograd to < 2e-3 per tensor (bf16 operands: < 1.5e-2) -- the bound the GPU test holds the real kernels to."""
    import logging
    from emulation import train_step_cpu_emulation
    from hific_b200.config import mse_lpips_args
    from hific_b200.grad import GRAD_BF16
    from hific_b200.model import Model
    from tools import grad_precision as GP
    tol = 1.5e-2 if GRAD_BF16 else 2e-3
    n_res = 1
    cfg = mse_lpips_args()
    cfg.n_residual_blocks = n_res
    sd2 = synth.synth_state_dict(0, n_residual_blocks=n_res)
    m = Model(cfg, logging.getLogger("forced-cpu"))
    m.load_state_dict(sd2, strict=True)
    m.train()
    g = torch.Generator().manual_seed(11)
    yh = torch.round(2 * torch.randn(1, 220, 8, 8, generator=g))
    upx = torch.randn(1, 3, 128, 128, generator=g)
    with train_step_cpu_emulation():
        def gen():
            yc = yh.clone().requires_grad_(True)
            xh = m.Generator(yc)
            plan = m.Generator._train_plans.get(yc)
            zs = {"init": GP.rows_to_nchw(plan.z_init, 1, 8, 8, 960)}
            for k, (z1, z2) in enumerate(plan.zr):
                zs[("r", k, 0)], zs[("r", k, 1)] = GP.rows_to_nchw(z1, 1, 8, 8, 960), GP.rows_to_nchw(z2, 1, 8, 8, 960)
            for i, (z, lay) in enumerate(zip(plan.zu, plan.ups)):
                zs[("u", i + 1)] = GP.rows_to_nchw(z, 1, lay.oh, lay.ow, lay.cout)
            (xh * upx).sum().backward()
            return xh.detach(), yc.grad, zs
        r = GP.study("Generator", m.Generator, "Generator.", sd2, gen,
                     lambda s, t, rr: O.generator_forward(s, t, n_residual_blocks=n_res, rnd=rr), [yh], upx,
                     forced=lambda s, t, zs, rr: GP.forced_generator(s, t, n_res, zs, rr))
    f = r["forced"]
    assert f["forward_rel_l2"] < 1e-4 and f["worst_tensor_rel_l2"] < tol and f["input_grad_rel_l2"] < tol, f
    assert r["fp32"]["worst_tensor_rel_l2"] > 3 * f["worst_tensor_rel_l2"]      # the flips, not the arithmetic, dominate


def test_backward_against_an_overwritten_plan_raises():
    """Saved activations live on the cached plan of an input shape: a second forward of the same shape before
    the first one's backward must make that backward FAIL, not differentiate through the wrong activations."""
    import pytest
    from emulation import training_cpu_emulation
    from hific_b200.network import hyper
    torch.manual_seed(0)
    net = hyper.HyperpriorAnalysis(C=220, N=320).train()
    g = torch.Generator().manual_seed(1)
    y1 = torch.randn((1, 220, 8, 8), generator=g).requires_grad_(True)
    y2 = torch.randn((1, 220, 8, 8), generator=g).requires_grad_(True)
    with training_cpu_emulation():
        z1 = net(y1)
        z2 = net(y2)                       # overwrites the plan's saved tensors
        z2.sum().backward()                # the latest forward is fine
        with pytest.raises(RuntimeError, match="overwritten by a later forward"):
            z1.sum().backward()
        # step by step is fine
        for p in net.parameters():
            p.grad = None
        net(y1).sum().backward()
        net(y2).sum().backward()
        assert all(p.grad is not None for p in net.parameters())
