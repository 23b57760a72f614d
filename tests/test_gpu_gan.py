"""GPU parity of the discriminator / loss half of the training step (forward) against the CPU oracle and the golden
vector produced by the REAL reference in COMPRESSION_GAN training mode (oracle/make_golden.py: gan_train_128)."""
import logging
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from hific_b200 import ops, synth  # noqa: E402
from hific_b200.config import ModelModes, ModelTypes, hific_args  # noqa: E402
from hific_b200.model import Model  # noqa: E402
from oracle import hific_oracle as O  # noqa: E402
from test_gpu_parity import Feed, rel_l2  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sd():
    return synth.synth_state_dict(0, gan=True)


@pytest.fixture()
def model(sd):
    cfg = hific_args()
    cfg.latent_dims = (220, 8, 8)
    m = Model(cfg, logging.getLogger("gan"), model_type=ModelTypes.COMPRESSION_GAN)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def test_discriminator_forward_and_spectral_norm(model, sd):
    """Fed the oracle's tensors: logits within the fp16-operand tolerance; u / v buffers updated like torch's hook."""
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4, 3, 128, 128, generator=g)
    y = torch.round(2 * torch.randn(4, 220, 8, 8, generator=g))
    for training in (True, False):
        sd_local = {k: v.clone() for k, v in sd.items()}
        model.load_state_dict(sd_local, strict=True)
        model.train(training)
        with torch.no_grad():
            ref_out, ref_logits, new_uv = O.discriminator_forward(sd_local, x, y, training=training)
            out, logits = model.Discriminator(x.cuda(), y.cuda())
        assert logits.shape == ref_logits.shape == (4 * 64, 1)
        assert rel_l2(logits, ref_logits) < 1e-3
        assert torch.allclose(out.cpu(), ref_out, atol=2e-3)
        for i in range(1, 5):
            u = getattr(model.Discriminator, f"conv{i}").weight_u.cpu()
            if training:
                assert torch.allclose(u, new_uv[f"conv{i}"][0], atol=1e-5), f"conv{i}.weight_u after power iteration"
            else:
                assert torch.equal(u, sd[f"Discriminator.conv{i}.weight_u"])


def test_loss_kernels_against_torch():
    g = torch.Generator().manual_seed(12)
    a, b = torch.rand(3, 3, 64, 48, generator=g).cuda(), torch.rand(3, 3, 64, 48, generator=g).cuda()
    ref = torch.mean((a * 255. - b * 255.) ** 2)
    assert abs(float(ops.sqdiff_sum(a, b, 255.)) / a.numel() - float(ref)) <= 1e-5 * float(ref)
    lg = (3 * torch.randn(2 * 500, generator=g)).cuda()
    s = ops.gan_sums(lg).float() / 500
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    real, gen = lg[:500], lg[500:]
    assert torch.allclose(s[0], bce(real, torch.ones_like(real)), rtol=1e-5)
    assert torch.allclose(s[1], bce(gen, torch.zeros_like(gen)), rtol=1e-5)
    assert torch.allclose(s[2], bce(gen, torch.ones_like(gen)), rtol=1e-5)
    assert torch.allclose(s[3], torch.sigmoid(real).mean(), rtol=1e-5)


def test_lpips_against_oracle(model):
    g = torch.Generator().manual_seed(13)
    pred, target = torch.rand(2, 3, 128, 128, generator=g), torch.rand(2, 3, 128, 128, generator=g)
    pl = model.perceptual_loss
    assert pl.lin_source.endswith("lpips_alex_lin_v0.1.npz")
    with torch.no_grad():
        got = pl.forward(pred.cuda(), target.cuda(), normalize=True).cpu()
        trunk_cpu = torch.nn.Sequential(*[m for m in pl.trunk]).cpu().float()
        ref = O.lpips_forward(trunk_cpu, [l.detach().cpu() for l in pl.lins], pred, target)
        pl.trunk.cuda()
    assert got.shape == ref.shape == (2, 1, 1, 1)
    assert torch.allclose(got, ref, rtol=5e-3, atol=1e-5)     # trunk convs run on cuDNN (TF32 allowed by torch default)


def test_training_forward_matches_reference_golden(model, sd):
    """Model.forward(train_generator=True) in COMPRESSION_GAN training mode vs the real reference's numbers."""
    gold = np.load(os.path.join(GOLDEN, "gan_train_128.npz"))
    x = synth.synth_image(2, 128, 128, 0)
    nz = synth.synth_noise((2, 320, 2, 2), "zgan", 0)
    ny = synth.synth_noise((2, 220, 8, 8), "ygan", 0)
    model.train(True)
    with torch.no_grad(), Feed([nz, ny]):
        losses, inter = model(x.cuda(), train_generator=True, return_intermediates=True)
    assert set(losses) == {"compression", "disc"}
    assert abs(float(inter.n_bpp) - float(gold["n_bpp"])) <= 2e-3 * float(gold["n_bpp"])
    assert abs(float(inter.q_bpp) - float(gold["q_bpp"])) <= 5e-3 * float(gold["q_bpp"])
    # the image (and everything computed from it) carries the y_hat rounding flips of the fp16-operand encoder, which
    # the randomly initialised generator amplifies: loose bounds here, strict stage-wise checks elsewhere
    step = int(gold["recon.stride"]) if "recon.stride" in gold else 1
    sub = gold["recon.sub"] if "recon.sub" in gold else gold["recon.full"].reshape(-1)
    got = inter.reconstruction.cpu().numpy().reshape(-1)[::step]
    assert np.linalg.norm(got - sub) / np.linalg.norm(sub) < 0.3
    assert abs(float(losses["disc"]) - float(gold["disc_loss"])) <= 0.05 * abs(float(gold["disc_loss"]))
    assert abs(float(losses["compression"]) - float(gold["compression_loss"])) <= 0.1 * abs(float(gold["compression_loss"]))
    assert model.step_counter == 1


def test_losses_given_oracle_reconstruction(model, sd):
    """Strict: feed the loss half the ORACLE's intermediates so no rounding flip is involved."""
    from hific_b200.model import Intermediates
    x = synth.synth_image(2, 128, 128, 0)
    nz = synth.synth_noise((2, 320, 2, 2), "zgan", 0)
    ny = synth.synth_noise((2, 220, 8, 8), "ygan", 0)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        recon, hyper, _ = O.compression_forward(sd, x, True, False, nz, ny)
        d_in = torch.cat([x, recon], 0)
        lat = torch.repeat_interleave(hyper.decoded, 2, dim=0)
        _, logits, _ = O.discriminator_forward(sd, d_in, lat, training=True)
        d_real, d_gen = torch.chunk(logits.squeeze(), 2, dim=0)
        d_loss_o, g_loss_o = O.gan_losses_non_saturating(d_real, d_gen)
        dist_o = O.distortion_loss(recon, x)
    model.train(True)
    inter = Intermediates(x.cuda(), recon.cuda(), hyper.decoded.cuda(), hyper.total_nbpp.cuda(), hyper.total_qbpp.cuda())
    model.writeout = False
    with torch.no_grad():
        d_loss, g_loss = model.GAN_loss(inter, train_generator=True)
        dist = model.distortion_loss(inter.reconstruction, inter.input_image)
    assert abs(float(d_loss) - float(d_loss_o)) <= 2e-3 * abs(float(d_loss_o))
    assert abs(float(g_loss) - float(g_loss_o)) <= 2e-3 * abs(float(g_loss_o))
    assert abs(float(dist) - float(dist_o)) <= 1e-5 * float(dist_o)


# ------------------------------------------------------------------------------------------------------------
# backward of the discriminator / GAN losses
# ------------------------------------------------------------------------------------------------------------
GRAD_TOL = 5e-2     # bf16-operand backward GEMMs on top of the fp16-operand forward (see tests/test_gpu_train.py)
D_PARAMS = (["context_conv.weight", "context_conv.bias"] + [f"conv{i}.{s}" for i in range(1, 5) for s in ("weight_orig", "bias")] +
            ["conv_out.weight", "conv_out.bias"])


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_gan_grad_kernel_against_torch():
    g = torch.Generator().manual_seed(21)
    for mode in (0, 1):
        real = (2 * torch.randn(300, generator=g)).cuda().requires_grad_(True)
        gen = (2 * torch.randn(300, generator=g)).cuda().requires_grad_(True)
        loss = ops.GanLossFn.apply(real, gen, mode)
        (3.0 * loss).backward()
        r2, g2 = real.detach().clone().requires_grad_(True), gen.detach().clone().requires_grad_(True)
        bce = torch.nn.functional.binary_cross_entropy_with_logits
        ref = bce(g2, torch.ones_like(g2)) if mode == 0 else bce(r2, torch.ones_like(r2)) + bce(g2, torch.zeros_like(g2))
        (3.0 * ref).backward()
        assert torch.allclose(loss, ref, rtol=1e-5)
        assert torch.allclose(gen.grad, g2.grad, rtol=1e-4, atol=1e-8)
        if mode == 1:
            assert torch.allclose(real.grad, r2.grad, rtol=1e-4, atol=1e-8)
        else:
            assert real.grad is None or float(real.grad.abs().max()) == 0.0


def test_discriminator_backward_vs_oracle(model, sd):
    """All 12 parameter gradients (incl. the spectral-norm reparametrisation) and d/dx of the discriminator loss
    against CPU autograd of the oracle (whose backward is pinned to the real reference by oracle/make_golden.py)."""
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(31)
    x = torch.rand(4, 3, 128, 128, generator=g)
    y = torch.round(2 * torch.randn(4, 220, 8, 8, generator=g))
    sd_local = {k: v.clone() for k, v in sd.items()}
    model.load_state_dict(sd_local, strict=True)
    model.train(True)
    xc = x.cuda().requires_grad_(True)
    _, logits = model.Discriminator(xc, y.cuda())
    d_real, d_gen = torch.chunk(logits.squeeze(), 2, dim=0)
    loss = ops.GanLossFn.apply(d_real, d_gen, 1) + 0.5 * ops.GanLossFn.apply(d_real, d_gen, 0)
    loss.backward()

    sdg = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("weight_u", "weight_v")))
           for k, v in sd.items() if k.startswith("Discriminator.")}
    xo = x.clone().requires_grad_(True)
    _, lo, _ = O.discriminator_forward(sdg, xo, y, training=True)
    r_o, g_o = torch.chunk(lo.squeeze(), 2, dim=0)
    dl, gl = O.gan_losses_non_saturating(r_o, g_o)
    (dl + 0.5 * gl).backward()
    assert abs(float(loss) - float(dl + 0.5 * gl)) < 2e-3 * abs(float(dl + 0.5 * gl))
    errs = {n: _rel(dict(model.Discriminator.named_parameters())[n].grad, sdg["Discriminator." + n].grad) for n in D_PARAMS}
    errs["x"] = _rel(xc.grad, xo.grad)
    print({k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < GRAD_TOL, errs


def test_discriminator_step_matches_reference_gradients(model, sd):
    """The D-step and the D-parameter part of the G-step vs the REAL reference's gradients (tests/golden/
    gan_grads_128.npz), fed the oracle's intermediates so that no y_hat rounding flip is involved."""
    from hific_b200.model import Intermediates
    gold = np.load(os.path.join(GOLDEN, "gan_grads_128.npz"))
    x = synth.synth_image(2, 128, 128, 0)
    nz = synth.synth_noise((2, 320, 2, 2), "zgan", 0)
    ny = synth.synth_noise((2, 220, 8, 8), "ygan", 0)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        recon, hyper, _ = O.compression_forward(sd, x, True, False, nz, ny)
    for tag, train_generator, scale in (("dstep", False, 1.0), ("gstep", True, 0.15)):
        model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        model.train(True)
        model.zero_grad(set_to_none=True)
        inter = Intermediates(x.cuda(), recon.cuda().requires_grad_(train_generator), hyper.decoded.cuda(),
                              hyper.total_nbpp.cuda(), hyper.total_qbpp.cuda())
        d_loss, g_loss = model.GAN_loss(inter, train_generator=train_generator)
        (d_loss if not train_generator else scale * g_loss).backward()
        for n in D_PARAMS:
            p = dict(model.Discriminator.named_parameters())[n]
            key = f"{tag}.Discriminator.{n}"
            norm = float(gold[key + ".norm"])
            flat = p.grad.reshape(-1).cpu()
            sub = flat[:: max(1, flat.numel() // 64)][:64].numpy()
            assert abs(float(p.grad.norm()) - norm) < GRAD_TOL * norm, (key, float(p.grad.norm()), norm)
            assert np.linalg.norm(sub - gold[key + ".sub"]) < 2 * GRAD_TOL * np.linalg.norm(gold[key + ".sub"]) + 1e-3 * norm / 8, key
