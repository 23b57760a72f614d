"""Golden vectors of the InstanceNorm2d variant (use_channel_norm = False) from the REAL reference modules
(/root/reference/src/network/encoder.py, generator.py with channel_norm=False) -- TEST INFRASTRUCTURE.

    python oracle/make_golden_instance.py        # build container only; writes tests/golden/instance_norm.npz

Inputs and weights are regenerated from seeds (hific_b200.synth); the fixture holds the reference's outputs and the
gradients of sum(out * wgt) w.r.t. every parameter (strided subsets + moments), forward in train() mode (InstanceNorm2d
without running statistics behaves the same in eval()).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import GOLDEN_DIR, summarize  # noqa: E402

SEED = 3
N_RES = 2


def inputs():
    from hific_b200 import synth
    g = torch.Generator().manual_seed(11)
    x = synth.synth_image(2, 64, 64, SEED)
    w_enc = torch.randn((2, 220, 4, 4), generator=g)
    y_hat = torch.round(torch.randn((2, 220, 4, 6), generator=g) * 2)
    w_gen = torch.randn((2, 3, 64, 96), generator=g)
    return x, w_enc, y_hat, w_gen


def state_dicts():
    from hific_b200 import synth
    sd = synth.instance_norm_variant(synth.synth_state_dict(SEED, n_residual_blocks=N_RES))
    enc = {k[len("Encoder."):]: v for k, v in sd.items() if k.startswith("Encoder.")}
    gen = {k[len("Generator."):]: v for k, v in sd.items() if k.startswith("Generator.")}
    return sd, enc, gen


def main():
    ref_shim.install()
    from src.network import encoder as ref_encoder, generator as ref_generator
    x, w_enc, y_hat, w_gen = inputs()
    _, enc_sd, gen_sd = state_dicts()
    out = {}
    enc = ref_encoder.Encoder((3, 64, 64), 2, C=220, channel_norm=False)
    enc.load_state_dict(enc_sd, strict=True)          # the keys / shapes of the variant are the reference's
    enc.train()
    y = enc(x)
    (y * w_enc).sum().backward()
    summarize("enc.y", y, out)
    for name, p in enc.named_parameters():
        summarize("enc.grad." + name, p.grad, out, full_limit=3000)
    gen = ref_generator.Generator((220, 4, 6), 2, C=220, n_residual_blocks=N_RES, channel_norm=False)
    gen.load_state_dict(gen_sd, strict=True)
    gen.train()
    yh = y_hat.clone().requires_grad_(True)
    xh = gen(yh)
    (xh * w_gen).sum().backward()
    summarize("gen.x_hat", xh, out)
    summarize("gen.grad.input", yh.grad, out, full_limit=3000)
    for name, p in gen.named_parameters():
        summarize("gen.grad." + name, p.grad, out, full_limit=3000)
    path = os.path.join(GOLDEN_DIR, "instance_norm.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
