"""CPU: replay the committed golden vectors (produced by the REAL reference, oracle/make_golden.py) against
the oracle restatement.  This is what pins the oracle; the GPU parity tests then compare CUDA against it."""
import os

import numpy as np
import pytest
import torch

from hific_b200 import synth
from oracle import hific_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "train_128": (2, 128, 128, True, False),
    "eval_128": (2, 128, 128, False, False),
    "train_256": (1, 256, 256, True, False),
    "evalmode_100x144": (1, 100, 144, False, True),
}


def check_summary(gold, name, t, atol):
    a = t.detach().numpy().astype(np.float32)
    assert tuple(gold[name + ".shape"]) == a.shape
    if name + ".full" in gold:
        np.testing.assert_allclose(a, gold[name + ".full"], rtol=0, atol=atol)
    else:
        step = int(gold[name + ".stride"])
        np.testing.assert_allclose(a.reshape(-1)[::step], gold[name + ".sub"], rtol=0, atol=atol)
    s, q = a.astype(np.float64).sum(), (a.astype(np.float64) ** 2).sum()
    assert abs(s - float(gold[name + ".sum"])) <= 1e-5 * max(1.0, abs(float(gold[name + ".sqsum"])) ** 0.5 * a.size ** 0.5)
    assert abs(q - float(gold[name + ".sqsum"])) <= 1e-5 * float(gold[name + ".sqsum"])


@pytest.fixture(scope="module")
def sd():
    return synth.synth_state_dict(0)


def golden_inputs(name):
    b, h, w, training, evalmode = CASES[name]
    x = synth.synth_image(b, h, w, 0)
    hp, wp = (-(-h // 16) * 16, -(-w // 16) * 16) if evalmode else (h, w)
    yh, yw = hp // 16, wp // 16
    if evalmode:
        yh, yw = -(-yh // 4) * 4, -(-yw // 4) * 4
    noise_z = synth.synth_noise((b, 320, yh // 4, yw // 4), f"z{name}", 0)
    noise_y = synth.synth_noise((b, 220, yh, yw), f"y{name}", 0)
    return x, noise_z, noise_y, training, evalmode


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name, sd):
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    x, noise_z, noise_y, training, evalmode = golden_inputs(name)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        recon, hyper, y = O.compression_forward(sd, x, training, evalmode, noise_z, noise_y)
    if not evalmode:
        check_summary(gold, "y", y, 1e-5)
    check_summary(gold, "decoded", hyper.decoded, 1e-5)
    check_summary(gold, "recon", recon, 1e-5)
    for f in ("latent_nbpp", "hyperlatent_nbpp", "total_nbpp", "latent_qbpp", "hyperlatent_qbpp", "total_qbpp"):
        assert abs(float(getattr(hyper, f)) - float(gold[f])) <= 1e-5 * max(1.0, abs(float(gold[f]))), f


def test_state_dict_contract_shapes():
    shapes = synth.hific_shapes()
    assert len(shapes) == 148
    assert shapes["Generator.upconv_block1.0.weight"] == (960, 480, 3, 3)
    assert shapes["Hyperprior.hyperlatent_likelihood.H_1"] == (320, 3, 3)
    total = sum(int(np.prod(s)) for s in shapes.values())
    assert total == 7423420 + 156774683 + 17277560  # SURVEY.md section 8a parameter counts
    gan = synth.hific_shapes(gan=True)
    # weight_u / weight_v are buffers, not parameters: 2 793 117 trainable discriminator parameters
    extra = sum(int(np.prod(s)) for k, s in gan.items()
                if k.startswith("Discriminator") and not k.endswith(("weight_u", "weight_v")))
    assert extra == 2793117


def test_lower_bound_and_quantise_edge_cases():
    x = torch.tensor([-0.5, 0.5, 1.5, -1.5, 0.49999, 2.5])
    mu = torch.zeros_like(x)
    assert torch.equal(O.quantize_st(x, mu), torch.floor(x + 0.5))
    lik = O.latent_likelihood(torch.tensor([100.0]), torch.tensor([0.0]), torch.tensor([0.11]))
    assert lik.item() == pytest.approx(1e-9)
    assert O.lower_bound(torch.tensor([0.05, 0.2]), 0.11).tolist() == pytest.approx([0.11, 0.2])


def test_oracle_backward_matches_reference_gradients():
    """The oracle's autograd (LowerBoundToward gates, straight-through quantisation, spectral-norm constants, the
    alternating G / D steps of train.py:137-141) against the parameter gradients of the REAL reference stored in
    tests/golden/gan_grads_128.npz (written by oracle/make_golden.py, where the full tensors agreed bit-exactly)."""
    import torchvision
    gold = np.load(os.path.join(GOLDEN, "gan_grads_128.npz"))
    sd = synth.synth_state_dict(0, gan=True)
    x = synth.synth_image(2, 128, 128, 0)
    nz = synth.synth_noise((2, 320, 2, 2), "zgan", 0)
    ny = synth.synth_noise((2, 220, 8, 8), "ygan", 0)
    state = torch.random.get_rng_state()
    torch.manual_seed(1234)                               # the offline AlexNet stand-in of oracle/ref_shim.py
    feats = torchvision.models.alexnet(weights=None).features.eval()
    torch.random.set_rng_state(state)
    for p in feats.parameters():
        p.requires_grad_(False)
    lin = np.load(os.path.join(os.path.dirname(synth.__file__), "weights", "lpips_alex_lin_v0.1.npz"))
    lins = [torch.from_numpy(lin[f"lin{k}"]) for k in range(5)]
    cfg = dict(lambda_A=2 ** 1, lambda_B=2 ** (-4), target_rate=0.14, lambda_schedule=dict(vals=[2., 1.], steps=[50000]),
               target_schedule=dict(vals=[0.20 / 0.14, 1.], steps=[50000]), k_M=0.075 * 2 ** (-5), k_P=1.0, beta=0.15)
    torch.set_num_threads(os.cpu_count())
    for tag, train_generator in (("dstep", False), ("gstep", True)):
        sdg = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("weight_u", "weight_v")))
               for k, v in sd.items()}
        comp, d_loss, _ = O.gan_training_losses(sdg, x, nz, ny, cfg, feats, lins, train_generator, step=1)
        (comp if train_generator else d_loss).backward()
        keys = [k[len(tag) + 1:-len(".norm")] for k in gold.files if k.startswith(tag + ".") and k.endswith(".norm")]
        assert len(keys) == (160 if train_generator else 12)
        for k in keys:
            g = sdg[k].grad
            assert g is not None, k
            norm = float(gold[f"{tag}.{k}.norm"])
            flat = g.reshape(-1)
            sub = flat[:: max(1, flat.numel() // 64)][:64].numpy()
            assert abs(float(g.norm()) - norm) <= 1e-4 * norm + 1e-12, (tag, k)
            assert np.allclose(sub, gold[f"{tag}.{k}.sub"], rtol=1e-3, atol=1e-5 * norm + 1e-12), (tag, k)
