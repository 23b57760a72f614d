"""Generate tests/golden/dlmm_c8.npz from the REAL reference's HyperpriorDLMM (src/hyperprior.py:340-458, imported from
/root/reference) -- TEST INFRASTRUCTURE, build container only.

    python oracle/make_golden_dlmm.py

The module is built under torch.manual_seed(SEED) with bottleneck_capacity 8 (its weights -- 5.8 M floats -- are NOT stored:
the product's mirror module reproduces them bit for bit under the same seed, which the tests check through the outputs).
Stored: the latents, the two noise tensors, every HyperInfo field in train and eval mode, the mixture parameters, and the
gradients of (total_nbpp) w.r.t. the latents and a few parameter tensors.  The script ends by checking the oracle
restatement (oracle/hific_oracle.py: hyperprior_dlmm_forward) against everything it wrote.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

SEED, C, SHAPE = 21, 8, (2, 8, 16, 16)
GRAD_KEYS = ("analysis_net.conv1.weight", "synthesis_DLMM_params.conv_out.weight", "synthesis_DLMM_params.conv3.bias",
             "synthesis_DLMM_params.conv3.weight", "hyperlatent_likelihood.H_1")


def inputs():
    g = torch.Generator().manual_seed(SEED + 1)
    y = torch.randn(SHAPE, generator=g) * 2.5
    nz = torch.rand((SHAPE[0], 320, SHAPE[2] // 4, SHAPE[3] // 4), generator=g) - 0.5
    ny = torch.rand(SHAPE, generator=g) - 0.5
    return y, nz, ny


def main():
    ref_shim.install()
    from src import hyperprior as ref_hp
    from oracle import hific_oracle as O
    torch.manual_seed(SEED)
    hp = ref_hp.HyperpriorDLMM(bottleneck_capacity=C)
    y, nz, ny = inputs()
    out = {"y": y.numpy(), "noise_z": nz.numpy(), "noise_y": ny.numpy()}
    fields = ("decoded", "latent_nbpp", "hyperlatent_nbpp", "total_nbpp", "latent_qbpp", "hyperlatent_qbpp", "total_qbpp")
    for training in (True, False):
        hp.train(training)
        tag = "train" if training else "eval"
        yy = y.clone().requires_grad_(True)
        for p in hp.parameters():
            p.grad = None
        with ref_shim.NoiseFeeder([nz, ny]):
            info = hp(yy, spatial_shape=(256, 256))
        for f in fields:
            out[f"{tag}.{f}"] = getattr(info, f).detach().numpy()
        if training:
            (info.total_nbpp * 1000.0 + info.decoded.square().mean()).backward()
            out["train.grad.y"] = yy.grad.numpy()
            named = dict(hp.named_parameters())
            for k in GRAD_KEYS:
                out["train.grad." + k] = named[k].grad.numpy()
    sd = {"Hyperprior." + k: v.detach() for k, v in hp.state_dict().items()}
    with torch.no_grad():
        z = O.hyper_analysis(sd, y)
        out["eval.dlmm_params"] = hp.synthesis_DLMM_params(torch.floor(z + 0.5)).numpy()
    path = os.path.join(ROOT, "tests", "golden", "dlmm_c8.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

    # ---- oracle vs reference (bit-exact forward; gradients to float tolerance of a different op order: none expected)
    g = np.load(path)
    for training in (True, False):
        tag = "train" if training else "eval"
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        yy = y.clone().requires_grad_(True)
        o = O.hyperprior_dlmm_forward(sdg, yy, (256, 256), training, nz, ny)
        for f in fields:
            a, b = getattr(o, f).detach().numpy(), g[f"{tag}.{f}"]
            assert np.array_equal(a, b), (tag, f, np.abs(a - b).max())
        if training:
            (o.total_nbpp * 1000.0 + o.decoded.square().mean()).backward()
            assert np.array_equal(yy.grad.numpy(), g["train.grad.y"])
            for k in GRAD_KEYS:
                assert np.array_equal(sdg["Hyperprior." + k].grad.numpy(), g["train.grad." + k]), k
        else:
            assert np.array_equal(o.latent_means.detach().numpy(), g["eval.dlmm_params"])
    print("oracle == reference HyperpriorDLMM (forward bit-exact in train and eval mode, gradients bit-exact)")


if __name__ == "__main__":
    main()
