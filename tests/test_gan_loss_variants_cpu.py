"""CPU: the least-squares GAN loss variant (`gan_loss_type='least_squares'`, src/loss/losses.py:43-50, 52-66) of
hific_b200.loss.losses.gan_loss -- values and gradients against the formulas of the reference, and against the reference's
own function where /root/reference is available (build container)."""
from collections import namedtuple

import pytest
import torch

from hific_b200.loss import losses as L
from oracle import ref_shim

Disc_out = namedtuple("Disc_out", ["D_real", "D_gen", "D_real_logits", "D_gen_logits"])


def _disc_out(seed):
    g = torch.Generator().manual_seed(seed)
    lr = torch.randn((2 * 256, 1), generator=g).requires_grad_(True)
    lg = torch.randn((2 * 256, 1), generator=g).requires_grad_(True)
    return Disc_out(torch.sigmoid(lr), torch.sigmoid(lg), lr, lg), lr, lg


@pytest.mark.parametrize("mode", ["generator_loss", "discriminator_loss"])
def test_least_squares_gan_loss_formula(mode):
    out, lr, lg = _disc_out(0)
    loss = L.gan_loss("least_squares", out, mode)
    if mode == "generator_loss":
        want = 0.5 * ((out.D_gen - 1.0) ** 2).mean()
    else:
        want = 0.5 * (((out.D_real - 1.0) ** 2).mean() + (out.D_gen ** 2).mean())
    assert torch.allclose(loss, want, rtol=0, atol=0)
    loss.backward()
    assert lg.grad is not None and (mode == "generator_loss") == (lr.grad is None)


def test_invalid_gan_loss_type_raises_like_the_reference():
    out, _, _ = _disc_out(1)
    with pytest.raises(ValueError):
        L.gan_loss("hinge", out, "generator_loss")


@pytest.mark.skipif(not ref_shim.available(), reason="needs the reference checkout (/root/reference)")
@pytest.mark.parametrize("mode", ["generator_loss", "discriminator_loss"])
def test_least_squares_gan_loss_equals_the_references(mode):
    ref_shim.install()
    from src.loss import losses as R
    out, lr, lg = _disc_out(2)
    ours = L.gan_loss("least_squares", out, mode)
    ours.backward()
    g_ours = (None if lr.grad is None else lr.grad.clone(), lg.grad.clone())
    lr.grad = lg.grad = None
    out2 = Disc_out(torch.sigmoid(lr), torch.sigmoid(lg), lr, lg)
    theirs = R.gan_loss("least_squares", out2, mode)
    theirs.backward()
    assert float(ours) == float(theirs)
    assert torch.equal(g_ours[1], lg.grad) and ((g_ours[0] is None and lr.grad is None) or torch.equal(g_ours[0], lr.grad))
