"""CPU: the inference PLANS of hific_b200.engine (layer lists, materialised reflect borders, asymmetric pads, stride-2 and
transposed layers, channel padding, fused / stand-alone ChannelNorm, the residual trunk) against the oracle, with the
activation-format entry points replaced by torch stand-ins (tests/emulation.py).  This checks host logic only -- which
buffers, geometries and parameters each launch gets; the kernels themselves are checked by the `-m gpu` tests.
Tolerance: fp16 storage of every intermediate activation -> 3e-3 relative L2 per network (the GPU tests hold the real
kernels to 1e-3 per stage)."""
import pytest
import torch

from emulation import plan_cpu_emulation
from hific_b200 import synth
from hific_b200.network import encoder, generator, hyper
from oracle import hific_oracle as O

REL = 3e-3


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.fixture(scope="module")
def sd():
    return synth.synth_state_dict(0)


def load(module, sd, prefix):
    module.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, strict=True)
    return module.eval()


@pytest.mark.parametrize("n,h,w", [(1, 64, 64), (2, 48, 80)])
def test_encoder_plan(sd, n, h, w):
    enc = load(encoder.Encoder((3, h, w), n, C=220), sd, "Encoder.")
    x = synth.synth_image(n, h, w, 3)
    with torch.no_grad(), plan_cpu_emulation():
        y = enc(x)
    want = O.encoder_forward(sd, x)
    assert y.shape == want.shape == (n, 220, h // 16, w // 16)
    assert rel_l2(y, want) < REL


@pytest.mark.parametrize("n,h,w", [(1, 4, 4), (2, 3, 5)])
def test_generator_plan(sd, n, h, w):
    gen = load(generator.Generator((220, h, w), n, C=220, n_residual_blocks=9), sd, "Generator.")
    g = torch.Generator().manual_seed(5)
    y_hat = torch.round(torch.randn((n, 220, h, w), generator=g) * 2)
    with torch.no_grad(), plan_cpu_emulation():
        x_hat = gen(y_hat)
    want = O.generator_forward(sd, y_hat)
    assert x_hat.shape == want.shape == (n, 3, 16 * h, 16 * w)
    assert rel_l2(x_hat, want) < REL


@pytest.mark.parametrize("n,h,w", [(1, 16, 16), (2, 8, 12)])
def test_hyper_plans(sd, n, h, w):
    ana = load(hyper.HyperpriorAnalysis(C=220, N=320), sd, "Hyperprior.analysis_net.")
    syn = load(hyper.HyperpriorSynthesis(C=220, N=320), sd, "Hyperprior.synthesis_mu.")
    g = torch.Generator().manual_seed(7)
    y = torch.randn((n, 220, h, w), generator=g)
    with torch.no_grad(), plan_cpu_emulation():
        z = ana(y)
        mu = syn(torch.round(z))
    want_z = O.hyper_analysis(sd, y)
    assert z.shape == want_z.shape == (n, 320, h // 4, w // 4) and rel_l2(z, want_z) < REL
    want_mu = O.hyper_synthesis(sd, torch.round(z), "Hyperprior.synthesis_mu.")
    assert mu.shape == want_mu.shape == (n, 220, h, w) and rel_l2(mu, want_mu) < REL


@pytest.mark.parametrize("b,h,w,evaluation", [(2, 128, 128, False), (1, 100, 144, True)])
def test_model_forward_eval_on_cpu(sd, b, h, w, evaluation):
    """`Model.forward` / `compression_forward` in eval mode through the emulated entry points: padding to multiples of
    16 / 4 and cropping (EVALUATION mode, ragged 100 x 144), hyperprior glue, bpp bookkeeping -- against the oracle."""
    import logging
    from emulation import model_cpu_emulation
    from hific_b200.config import ModelModes, mse_lpips_args
    from hific_b200.model import Model
    m = Model(mse_lpips_args(), logging.getLogger("cpu"),
              model_mode=ModelModes.EVALUATION if evaluation else ModelModes.TRAINING)
    assert not m.load_state_dict(sd, strict=False).unexpected_keys
    m.eval()
    x = synth.synth_image(b, h, w, 0)
    want_rec, want_hyp, _ = O.compression_forward(sd, x, training=False, evaluation_mode=evaluation)
    with torch.no_grad(), model_cpu_emulation():
        inter, info = m.compression_forward(x)
        if evaluation:
            rec, q_bpp = m(x, writeout=False)
            assert tuple(rec.shape) == (b, 3, h, w) and float(rec.min()) >= 0 and float(rec.max()) <= 1
            assert abs(float(q_bpp) - float(want_hyp.total_qbpp)) < 5e-3 * float(want_hyp.total_qbpp)
    assert tuple(inter.reconstruction.shape) == tuple(want_rec.shape)
    flips = ((inter.latents_quantized - want_hyp.decoded).abs() > 0.5).float().mean().item()
    # whole chain (encoder error reaches the rounding, unlike the stage-wise GPU protocol that feeds the oracle's y):
    # |y - mu| within ~1e-3 |y| of a rounding boundary flips -- a percent or two of the latents with random weights
    assert flips < 3e-2
    assert abs(float(info.total_qbpp) - float(want_hyp.total_qbpp)) < 5e-3 * float(want_hyp.total_qbpp)
    assert abs(float(info.hyperlatent_qbpp) - float(want_hyp.hyperlatent_qbpp)) < 5e-3 * float(want_hyp.hyperlatent_qbpp)
    # the generator amplifies the flipped latents: compare it on the oracle's latents instead (stage-wise protocol)
    with torch.no_grad(), model_cpu_emulation():
        xh = m.Generator(want_hyp.decoded)
    crop = xh[:, :, :want_rec.shape[2], :want_rec.shape[3]]
    assert rel_l2(crop, want_rec) < REL


@pytest.mark.parametrize("C", [64, 100])
def test_plans_with_other_latent_channels(C):
    """train.py exposes `--latent_channels` (train.py:240): the plans must follow any C, not only HiFIC's 220."""
    torch.manual_seed(C)
    enc = encoder.Encoder((3, 64, 64), 1, C=C).eval()
    gen = generator.Generator((C, 4, 4), 1, C=C, n_residual_blocks=2).eval()
    ana = hyper.HyperpriorAnalysis(C=C, N=320).eval()
    syn = hyper.HyperpriorSynthesis(C=C, N=320).eval()
    sd = {}
    for prefix, m in (("Encoder.", enc), ("Generator.", gen), ("Hyperprior.analysis_net.", ana),
                      ("Hyperprior.synthesis_mu.", syn)):
        sd.update({prefix + k: v.detach() for k, v in m.state_dict().items()})
    x = synth.synth_image(1, 64, 64, 1)
    g = torch.Generator().manual_seed(2)
    y16 = torch.randn((1, C, 16, 16), generator=g)
    with torch.no_grad(), plan_cpu_emulation():
        y = enc(x)
        x_hat = gen(torch.round(y))
        z = ana(y16)
        mu = syn(torch.round(z))
    assert rel_l2(y, O.encoder_forward(sd, x)) < REL
    assert rel_l2(x_hat, O.generator_forward(sd, torch.round(y), n_residual_blocks=2)) < REL
    assert rel_l2(z, O.hyper_analysis(sd, y16)) < REL
    assert rel_l2(mu, O.hyper_synthesis(sd, torch.round(z), "Hyperprior.synthesis_mu.")) < REL


def test_generator_plan_with_fused_residual_norms(sd, monkeypatch):
    """HFC_FUSE_RESNORM=1: the residual trunk through `Conv.call_widenorm` (conv + ChannelNorm + ReLU / residual adds in
    one launch) -- same result as the two-launch plan, i.e. as the oracle; ragged maps fall back to the two-launch plan."""
    from hific_b200 import engine
    monkeypatch.setenv("HFC_FUSE_RESNORM", "1")
    g = torch.Generator().manual_seed(5)
    for n, h, w, fused in ((2, 16, 16, True), (1, 3, 5, False)):
        gen = load(generator.Generator((220, h, w), n, C=220, n_residual_blocks=9), sd, "Generator.")
        y_hat = torch.round(torch.randn((n, 220, h, w), generator=g) * 2)
        with torch.no_grad(), plan_cpu_emulation():
            x_hat = gen(y_hat)
            plan = gen._plans.get(y_hat)
        assert (plan.fused is not None) == fused
        assert rel_l2(x_hat, O.generator_forward(sd, y_hat)) < REL


# ----------------------------------------------------------------------------------------------------------------------
# Non-default generator variants: sample_noise=True (992-channel trunk) and a stand-alone ResidualBlock
# ----------------------------------------------------------------------------------------------------------------------
class _FixedRandn:
    """The plans draw the generator noise with torch.randn as the reference does (generator.py:151): feed a known draw."""

    def __init__(self, noise):
        self.noise = noise

    def __enter__(self):
        self._orig = torch.randn
        torch.randn = lambda *a, **k: self.noise.clone()
        return self

    def __exit__(self, *e):
        torch.randn = self._orig


def _noise_generator(n_res=2):
    torch.manual_seed(21)
    gen = generator.Generator((220, 8, 8), 2, C=220, n_residual_blocks=n_res, sample_noise=True, noise_dim=32)
    with torch.no_grad():
        for name, p in gen.named_parameters():           # away from the identity initialisation of the norms
            if "gamma" in name:
                p.add_(0.1 * torch.randn(p.shape))
            elif "beta" in name or "bias" in name:
                p.add_(0.1 * torch.randn(p.shape))
    sdg = {"Generator." + k: v.detach().clone() for k, v in gen.state_dict().items()}
    return gen, sdg


def test_generator_with_sample_noise_inference_and_training_plans():
    from emulation import training_cpu_emulation
    gen, sdg = _noise_generator()
    g = torch.Generator().manual_seed(3)
    y_hat = torch.round(torch.randn((2, 220, 8, 8), generator=g) * 2)
    z = torch.randn((2, 32, 8, 8), generator=g)
    want = O.generator_forward(sdg, y_hat, n_residual_blocks=2, noise=z)
    gen.eval()
    with torch.no_grad(), plan_cpu_emulation(), _FixedRandn(z):
        got = gen(y_hat)
    assert gen._plans.get(y_hat).fused is None and gen._plans.get(y_hat).F0 == 992
    assert rel_l2(got, want) < REL
    # training plan: every parameter gradient and the input gradient against autograd of the oracle
    gen.train()
    up = torch.randn(want.shape, generator=g)
    yc = y_hat.clone().requires_grad_(True)
    with training_cpu_emulation(), _FixedRandn(z):
        out = gen(yc)
        (out * up).sum().backward()
    sd2 = {k: v.clone().requires_grad_(True) for k, v in sdg.items()}
    yo = y_hat.clone().requires_grad_(True)
    (O.generator_forward(sd2, yo, n_residual_blocks=2, noise=z) * up).sum().backward()
    assert rel_l2(out.detach(), want) < REL
    worst = max(rel_l2(p.grad, sd2["Generator." + k].grad) for k, p in gen.named_parameters())
    assert worst < 5e-2 and rel_l2(yc.grad, yo.grad) < 5e-2, worst
    assert gen.resblock_0.conv1.weight.shape == (992, 992, 3, 3) and gen.upconv_block1[0].weight.shape[0] == 992


def test_residual_block_called_on_its_own():
    from emulation import training_cpu_emulation
    torch.manual_seed(5)
    blk = generator.ResidualBlock((2, 128, 8, 8))
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(0.05 * torch.randn(p.shape))
    sdb = {"b." + k: v.detach().clone() for k, v in blk.state_dict().items()}
    g = torch.Generator().manual_seed(6)
    x = torch.randn((2, 128, 8, 8), generator=g)
    want = O.residual_block(sdb, "b", x)
    with torch.no_grad(), plan_cpu_emulation():
        blk.eval()
        got = blk(x)
    assert rel_l2(got, want) < REL
    up = torch.randn(want.shape, generator=g)
    xc = x.clone().requires_grad_(True)
    blk.train()
    with training_cpu_emulation():
        (blk(xc) * up).sum().backward()
    sd2 = {k: v.clone().requires_grad_(True) for k, v in sdb.items()}
    xo = x.clone().requires_grad_(True)
    (O.residual_block(sd2, "b", xo) * up).sum().backward()
    assert max(rel_l2(p.grad, sd2["b." + k].grad) for k, p in blk.named_parameters()) < 3e-2
    assert rel_l2(xc.grad, xo.grad) < 3e-2
