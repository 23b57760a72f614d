"""CPU: the InstanceNorm2d variant (use_channel_norm = False; src/normalisation/instance.py:7-15, encoder.py:41-44,
generator.py:21-24, 81-84).
  1. the oracle's restatement against golden vectors of the REAL reference modules built with channel_norm=False
     (tests/golden/instance_norm.npz, written by oracle/make_golden_instance.py): outputs and every parameter gradient;
  2. the product's Encoder / Generator / ResidualBlock with channel_norm=False -- module construction, the reference's
     state_dict names, the plans' layer walk with hfc_instancenorm / hfc_instancenorm_bwd in place of the ChannelNorm
     entry points (kernels replaced by the torch stand-ins of tests/emulation.py) -- against the oracle and its autograd.
The kernels themselves are checked on a GPU by tests/test_gpu_zzinstancenorm.py."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emulation import plan_cpu_emulation, training_cpu_emulation  # noqa: E402
from hific_b200.network import encoder, generator  # noqa: E402
from oracle import hific_oracle as O  # noqa: E402
from oracle.make_golden_instance import N_RES, inputs, state_dicts  # noqa: E402
from test_oracle_golden import check_summary  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "instance_norm.npz")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def grad_close(gold, name, t, rtol=2e-3):
    """Relative L2 on the stored subset (fp32 autograd of two mathematically equal graphs: summation orders differ)."""
    a = t.detach().numpy().astype(np.float64).reshape(-1)
    assert tuple(gold[name + ".shape"]) == tuple(t.shape), name
    ref = gold[name + ".full"].astype(np.float64).reshape(-1) if name + ".full" in gold else gold[name + ".sub"].astype(np.float64)
    got = a if name + ".full" in gold else a[::int(gold[name + ".stride"])]
    denom = max(np.linalg.norm(ref), 1e-6 * max(1.0, float(gold[name + ".sqsum"]) ** 0.5))
    assert np.linalg.norm(got - ref) / denom < rtol, name


def test_oracle_instance_variant_matches_reference_golden():
    gold = np.load(GOLD)
    x, w_enc, y_hat, w_gen = inputs()
    sd, _, _ = state_dicts()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(("Encoder.", "Generator."))}
    y = O.encoder_forward(sdg, x)
    (y * w_enc).sum().backward()
    check_summary(gold, "enc.y", y, atol=2e-4)
    yh = y_hat.clone().requires_grad_(True)
    xh = O.generator_forward(sdg, yh, n_residual_blocks=N_RES)
    (xh * w_gen).sum().backward()
    check_summary(gold, "gen.x_hat", xh, atol=2e-4)
    grad_close(gold, "gen.grad.input", yh.grad)
    for k, v in sdg.items():
        net, name = k.split(".", 1)
        if net == "Encoder" and name.startswith("conv_block") and name.endswith(".1.bias") and not name.startswith("conv_block_out"):
            # the bias of a conv in front of an InstanceNorm has a mathematically zero gradient (the norm removes the
            # per-channel mean): both sides hold rounding noise only
            assert float(v.grad.abs().max()) < 1e-3 * float(sdg[k.replace(".bias", ".weight")].grad.abs().max())
            continue
        if net == "Generator" and (name.endswith(("conv1.bias", "conv2.bias", "conv_block_init.2.bias")) or
                                   (name.startswith("upconv_block") and name.endswith(".0.bias"))):
            continue
        grad_close(gold, ("enc" if net == "Encoder" else "gen") + ".grad." + name, v.grad)


def test_state_dict_names_are_the_references():
    _, enc_sd, gen_sd = state_dicts()
    enc = encoder.Encoder((3, 64, 64), 2, C=220, channel_norm=False)
    gen = generator.Generator((220, 4, 6), 2, C=220, n_residual_blocks=N_RES, channel_norm=False)
    enc.load_state_dict(enc_sd, strict=True)      # the same dict the REAL reference modules loaded strictly (make_golden_instance)
    gen.load_state_dict(gen_sd, strict=True)
    assert isinstance(enc.conv_block1[2], torch.nn.InstanceNorm2d) and not enc.conv_block1[2].track_running_stats
    assert "conv_block1.2.weight" in enc.state_dict() and "resblock_0.norm1.bias" in gen.state_dict()


def _modules():
    sd, enc_sd, gen_sd = state_dicts()
    enc = encoder.Encoder((3, 64, 64), 2, C=220, channel_norm=False)
    gen = generator.Generator((220, 4, 6), 2, C=220, n_residual_blocks=N_RES, channel_norm=False)
    enc.load_state_dict(enc_sd, strict=True)
    gen.load_state_dict(gen_sd, strict=True)
    return sd, enc, gen


def test_inference_runs_the_instance_norm_plans():
    sd, enc, gen = _modules()
    x, _, y_hat, _ = inputs()
    enc.eval(), gen.eval()
    with torch.no_grad(), plan_cpu_emulation():
        y = enc(x)
        xh = gen(y_hat)
    assert rel(y, O.encoder_forward(sd, x)) < 3e-3
    assert rel(xh, O.generator_forward(sd, y_hat, n_residual_blocks=N_RES)) < 3e-3


def test_training_plans_against_oracle_autograd():
    sd, enc, gen = _modules()
    x, w_enc, y_hat, w_gen = inputs()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(("Encoder.", "Generator."))}
    (O.encoder_forward(sdg, x) * w_enc).sum().backward()
    yo = y_hat.clone().requires_grad_(True)
    (O.generator_forward(sdg, yo, n_residual_blocks=N_RES) * w_gen).sum().backward()
    enc.train(), gen.train()
    yp = y_hat.clone().requires_grad_(True)
    with training_cpu_emulation():
        (enc(x) * w_enc).sum().backward()
        (gen(yp) * w_gen).sum().backward()
    assert rel(yp.grad, yo.grad) < 8e-2
    for prefix, mod in (("Encoder.", enc), ("Generator.", gen)):
        for name, p in mod.named_parameters():
            want = sdg[prefix + name].grad
            if want.abs().max() < 1e-3 * sdg[prefix + name.replace(".bias", ".weight")].grad.abs().max() and name.endswith(".bias"):
                assert p.grad.abs().max() < 1e-2 * sdg[prefix + name.replace(".bias", ".weight")].grad.abs().max(), name
                continue                          # conv bias in front of an InstanceNorm: zero gradient on both sides
            assert rel(p.grad, want) < 8e-2, (prefix + name, rel(p.grad, want))


def test_residual_block_instance_variant_on_its_own():
    torch.manual_seed(5)
    blk = generator.ResidualBlock((2, 128, 8, 8), channel_norm=False)
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(0.05 * torch.randn(p.shape))
    sdb = {"b." + k: v.detach().clone() for k, v in blk.state_dict().items()}
    x = torch.randn((2, 128, 8, 8))
    want = O.residual_block(sdb, "b", x)
    blk.eval()
    with torch.no_grad(), plan_cpu_emulation():
        got = blk(x)
    assert rel(got, want) < 3e-3
