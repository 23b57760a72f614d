"""GPU: the `-LMM` variant -- csrc/dlmm.cu against the oracle / the closed-form stand-in of tests/emulation.py, and the
product's HyperpriorDLMM against the values of the REAL reference (tests/golden/dlmm_c8.npz).

First run on a B200 in round 2 (all cases passed as written).  Everything above the kernels is also covered on the CPU
by tests/test_dlmm_cpu.py.
Tolerances: sums 2e-5 relative (fp32 partial sums of ~1e4 terms), gradients 1e-4 relative L2 against the same formulas in
torch, module level 5e-3 / 5e-2 (fp16 activations, bf16 gradient operands) as for the other networks."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import emulation as E  # noqa: E402
from hific_b200 import hyperprior, ops  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dlmm_c8.npz")


def inputs(seed, n=2, c=5, k=4, h=6, w=7, spread=3.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, c, h, w), generator=g) * spread
    params = torch.randn((n, 3 * c * k, h, w), generator=g)
    params[:, 2 * c * k:] = params[:, 2 * c * k:] * 2 - 2.0
    params[:, c * k:2 * c * k] *= 3
    noise = torch.rand((n, c, h, w), generator=g) - 0.5
    return x, params, noise


@pytest.mark.parametrize("kind", ["gaussian", "logistic"])
@pytest.mark.parametrize("shape", [(2, 5, 4, 6, 7), (1, 64, 4, 16, 16), (3, 8, 2, 1, 3), (2, 16, 8, 5, 5)])
def test_forward_kernel(kind, shape):
    n, c, k, h, w = shape
    x, params, noise = inputs(sum(shape), n, c, k, h, w)
    for st in (True, False):
        want_dec, want = E.dlmm_likelihood(x, params, noise, kind, st)
        dec, sums = ops.dlmm_likelihood(x.cuda(), params.cuda(), noise.cuda(), kind, st)
        assert torch.equal(dec.cpu(), want_dec)
        for i in range(2):
            assert abs(sums[i].item() - want[i].item()) <= 2e-5 * abs(want[i].item()) + 1e-4
    dec, sums = ops.dlmm_likelihood(x.cuda(), params.cuda(), None, kind, False)
    assert sums[0].item() == 0.0


@pytest.mark.parametrize("kind", ["gaussian", "logistic"])
@pytest.mark.parametrize("upstream", [-0.37, 0.8])
def test_backward_kernel(kind, upstream):
    x, params, noise = inputs(11, spread=6.0)
    g = torch.Generator().manual_seed(1)
    dd = torch.randn(x.shape, generator=g)
    up = torch.tensor([upstream])
    want_dx, want_dp = E.dlmm_likelihood_bwd(x, params, noise, dd, up, kind)
    dx, dp = ops.dlmm_likelihood_bwd(x.cuda(), params.cuda(), noise.cuda(), dd.cuda(), up.cuda(), kind)
    assert ((dx.cpu() - want_dx).norm() / want_dx.norm()).item() < 1e-4
    assert ((dp.cpu() - want_dp).norm() / want_dp.norm()).item() < 1e-4
    dx2, _ = ops.dlmm_likelihood_bwd(x.cuda(), params.cuda(), noise.cuda(), None, up.cuda(), kind)
    assert torch.allclose(dx2.cpu(), want_dx - dd, rtol=1e-3, atol=1e-5 * float(want_dx.abs().max()))


def test_module_against_the_reference():
    gold = np.load(GOLDEN)
    fields = ("latent_nbpp", "hyperlatent_nbpp", "total_nbpp", "latent_qbpp", "hyperlatent_qbpp", "total_qbpp")
    keys = ("analysis_net.conv1.weight", "synthesis_DLMM_params.conv_out.weight", "synthesis_DLMM_params.conv3.bias",
            "synthesis_DLMM_params.conv3.weight", "hyperlatent_likelihood.H_1")
    y, nz, ny = (torch.from_numpy(gold[k]).cuda() for k in ("y", "noise_z", "noise_y"))
    torch.manual_seed(21)
    hp = hyperprior.HyperpriorDLMM(bottleneck_capacity=8).cuda().train()
    from oracle.ref_shim import NoiseFeeder
    yy = y.clone().requires_grad_(True)
    l0 = ops.launch_count()
    with NoiseFeeder([nz, ny]):
        info = hp(yy, spatial_shape=(256, 256))
    (info.total_nbpp * 1000.0 + info.decoded.square().mean()).backward()
    torch.cuda.synchronize()
    assert ops.launch_count() - l0 > 30
    rel = lambda a, b: ((a.cpu() - b).norm() / b.norm()).item()
    for f in fields:
        assert abs(float(getattr(info, f)) - float(gold[f"train.{f}"])) < 5e-3 * abs(float(gold[f"train.{f}"])), f
    assert torch.equal(info.decoded.detach().cpu(), torch.from_numpy(gold["train.decoded"]))
    assert rel(yy.grad, torch.from_numpy(gold["train.grad.y"])) < 5e-2
    grads = {k: v.grad for k, v in hp.named_parameters()}
    for k in keys:
        assert rel(grads[k], torch.from_numpy(gold["train.grad." + k])) < 5e-2, k
    hp.eval()
    with torch.no_grad(), NoiseFeeder([nz, ny]):
        info = hp(y, spatial_shape=(256, 256))
    for f in fields:
        assert abs(float(getattr(info, f)) - float(gold[f"eval.{f}"])) < 5e-3 * abs(float(gold[f"eval.{f}"])), f
    assert torch.equal(info.decoded.cpu(), torch.from_numpy(gold["eval.decoded"]))
