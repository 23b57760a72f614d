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
