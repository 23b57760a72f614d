"""GPU: schedules 2 (packed fp32, balanced persistent grid) and 3 (the default: 2 + register-double-buffered loads) of the
Gaussian latent-likelihood kernel (csrc/likelihood_v2.cu), selected with HFC_LIKELIHOOD_V, against a float64 restatement of src/hyperprior.py:124-139 and
against the first schedule on the same inputs.  Tolerance: 2e-5 relative on the log-likelihood sums (as for schedule 1),
straight-through latents bit-identical to schedule 1."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from hific_b200 import ops  # noqa: E402

SUM_RTOL = 2e-5


class Schedule:
    def __init__(self, v):
        self.v = str(v)

    def __enter__(self):
        self.old = os.environ.get("HFC_LIKELIHOOD_V")
        os.environ["HFC_LIKELIHOOD_V"] = self.v

    def __exit__(self, *e):
        if self.old is None:
            os.environ.pop("HFC_LIKELIHOOD_V", None)
        else:
            os.environ["HFC_LIKELIHOOD_V"] = self.old


def ref64(y, mu, sraw, noise, lb=0.11):
    sc = torch.clamp(sraw, min=lb).double()
    mu64 = mu.double()

    def loglik(x):
        d = (x.double() - mu64).abs()
        cdf = lambda v: 0.5 * torch.erfc(v * (-1.0 / math.sqrt(2)))
        p = torch.clamp(cdf((0.5 - d) / sc) - cdf(-(0.5 + d) / sc), min=1e-9)
        return torch.log(p + 1e-9).sum()

    v = y - mu
    vq = torch.floor(v + 0.5)
    dec = (v + (vq - v)) + mu                       # fp32, op order of quantize_latents_st
    sq = loglik((vq + mu))
    sn = loglik(y + noise) if noise is not None else None
    return dec, sn, sq


def inputs(shape, seed, scale=2.0):
    g = torch.Generator().manual_seed(seed)
    y = scale * torch.randn(shape, generator=g)
    mu = torch.randn(shape, generator=g)
    sraw = 2 * torch.rand(shape, generator=g) - 0.05       # some below the 0.11 bound, some negative
    noise = torch.rand(shape, generator=g) - 0.5
    return y, mu, sraw, noise


@pytest.mark.parametrize("shape", [(4, 220, 16, 16), (32, 220, 16, 16), (1, 3, 5, 7), (1, 1, 1, 2), (3, 7, 11, 13)])
@pytest.mark.parametrize("with_noise", [True, False])
@pytest.mark.parametrize("schedule", [2, 3, 4])
def test_schedule2_matches_float64_and_schedule1(shape, with_noise, schedule):
    y, mu, sraw, noise = inputs(shape, seed=sum(shape))
    if not with_noise:
        noise = None
    dec64, sn, sq = ref64(y, mu, sraw, noise)
    dev = [t.cuda() if t is not None else None for t in (y, mu, sraw, noise)]
    with Schedule(1):
        dec1, sums1 = ops.latent_likelihood(*dev, 0.11, "gaussian")
    l0 = ops.launch_count()
    with Schedule(schedule):
        dec2, sums2 = ops.latent_likelihood(*dev, 0.11, "gaussian")
    assert ops.launch_count() - l0 == 1
    torch.cuda.synchronize()
    assert torch.equal(dec2, dec1) and torch.equal(dec2.cpu(), dec64)
    floor = 1e-3 * y.numel()                                  # tiny tensors: absolute floor instead of a ratio
    assert abs(sums2[1].item() - sq.item()) <= SUM_RTOL * abs(sq.item()) + 1e-6 * floor
    assert abs(sums2[1].item() - sums1[1].item()) <= SUM_RTOL * abs(sq.item()) + 1e-6 * floor
    if with_noise:
        assert abs(sums2[0].item() - sn.item()) <= SUM_RTOL * abs(sn.item()) + 1e-6 * floor
    else:
        assert sums2[0].item() == 0.0


@pytest.mark.parametrize("schedule", [2, 3, 4])
def test_schedule2_far_tails_and_tiny_scales(schedule):
    """|y - mu| up to 1e4 at scale 0.11 (p underflows to the 1e-9 bound), large scales (p ~ 4e-3 from the difference of two values near 1), exact half-integers."""
    g = torch.Generator().manual_seed(3)
    n = 4096
    y = torch.cat([torch.randn(n, generator=g) * 1e4, torch.randn(n, generator=g) * 1e-3, torch.arange(n) * 0.5])
    mu = torch.zeros_like(y)
    sraw = torch.cat([torch.full((n,), 0.01), torch.full((n,), 1e2), torch.rand(n, generator=g) * 3])
    noise = torch.rand(y.shape, generator=g) - 0.5
    shape = (1, 3, 64, 64)
    y, mu, sraw, noise = (t.view(shape).contiguous() for t in (y, mu, sraw, noise))
    dec64, sn, sq = ref64(y, mu, sraw, noise)
    with Schedule(schedule):
        dec, sums = ops.latent_likelihood(y.cuda(), mu.cuda(), sraw.cuda(), noise.cuda(), 0.11, "gaussian")
    assert torch.equal(dec.cpu(), dec64)
    assert torch.isfinite(sums).all()
    assert abs(sums[0].item() - sn.item()) <= 2e-4 * abs(sn.item())
    assert abs(sums[1].item() - sq.item()) <= 2e-4 * abs(sq.item())


def test_logistic_stays_on_schedule1():
    y, mu, sraw, noise = (t.cuda() for t in inputs((2, 16, 8, 8), 9))
    with Schedule(1):
        d1, s1 = ops.latent_likelihood(y, mu, sraw, noise, 0.11, "logistic")
    with Schedule(2):
        d2, s2 = ops.latent_likelihood(y, mu, sraw, noise, 0.11, "logistic")
    assert torch.equal(d1, d2) and torch.allclose(s1, s2, rtol=1e-12)
