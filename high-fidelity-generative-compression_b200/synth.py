"""Deterministic synthetic parameters and inputs for tests and benchmarks.

There is no network in the build/benchmark environment (no datasets, no released checkpoints), so every
test and benchmark runs on procedurally generated weights: each tensor is drawn from its own CPU generator
seeded by crc32(key) ^ seed, so the same values can be regenerated anywhere (this container, the GPU box)
without shipping a 700 MB state dict.  Keys and shapes are the reference's ``state_dict`` contract
(SURVEY.md section 8b); ``oracle/make_golden.py`` asserts them against the real reference modules.
"""
import math
import zlib

import torch


def hific_shapes(C=220, n_residual_blocks=9, N=320, im_channels=3, gan=False):
    """key -> shape for Encoder / Generator / Hyperprior (/ Discriminator) in TRAINING mode."""
    s = {}

    def conv(prefix, cout, cin, k):
        s[prefix + ".weight"] = (cout, cin, k, k)
        s[prefix + ".bias"] = (cout,)

    def convT(prefix, cin, cout, k):
        s[prefix + ".weight"] = (cin, cout, k, k)
        s[prefix + ".bias"] = (cout,)

    def norm(prefix, c):
        s[prefix + ".gamma"] = (1, c, 1, 1)
        s[prefix + ".beta"] = (1, c, 1, 1)

    f = (60, 120, 240, 480, 960)
    conv("Encoder.conv_block1.1", f[0], im_channels, 7)
    norm("Encoder.conv_block1.2", f[0])
    for i in range(1, 5):
        conv(f"Encoder.conv_block{i + 1}.1", f[i], f[i - 1], 3)
        norm(f"Encoder.conv_block{i + 1}.2", f[i])
    conv("Encoder.conv_block_out.1", C, f[4], 3)

    norm("Generator.conv_block_init.0", C)
    conv("Generator.conv_block_init.2", 960, C, 3)
    norm("Generator.conv_block_init.3", 960)
    for m in range(n_residual_blocks):
        p = f"Generator.resblock_{m}"
        conv(p + ".conv1", 960, 960, 3)
        conv(p + ".conv2", 960, 960, 3)
        norm(p + ".norm1", 960)
        norm(p + ".norm2", 960)
    g = (960, 480, 240, 120, 60)
    for i in range(1, 5):
        convT(f"Generator.upconv_block{i}.0", g[i - 1], g[i], 3)
        norm(f"Generator.upconv_block{i}.1", g[i])
    conv("Generator.conv_block_out.1", im_channels, 60, 7)

    conv("Hyperprior.analysis_net.conv1", N, C, 3)
    conv("Hyperprior.analysis_net.conv2", N, N, 5)
    conv("Hyperprior.analysis_net.conv3", N, N, 5)
    for net in ("synthesis_mu", "synthesis_std"):
        convT(f"Hyperprior.{net}.conv1", N, N, 5)
        convT(f"Hyperprior.{net}.conv2", N, N, 5)
        convT(f"Hyperprior.{net}.conv3", N, C, 3)
    filt = (1, 3, 3, 3, 1)
    for k in range(4):
        s[f"Hyperprior.hyperlatent_likelihood.H_{k}"] = (N, filt[k + 1], filt[k])
        s[f"Hyperprior.hyperlatent_likelihood.a_{k}"] = (N, filt[k + 1], 1)
        s[f"Hyperprior.hyperlatent_likelihood.b_{k}"] = (N, filt[k + 1], 1)

    if gan:
        conv("Discriminator.context_conv", 12, C, 3)
        d = (im_channels + 12, 64, 128, 256, 512)
        for i in range(1, 5):
            s[f"Discriminator.conv{i}.bias"] = (d[i],)
            s[f"Discriminator.conv{i}.weight_orig"] = (d[i], d[i - 1], 4, 4)
            s[f"Discriminator.conv{i}.weight_u"] = (d[i],)
            s[f"Discriminator.conv{i}.weight_v"] = (d[i - 1] * 16,)
        conv("Discriminator.conv_out", 1, 512, 1)
    return s


def _gen(key, seed):
    return torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synth_tensor(key, shape, seed=0):
    g = _gen(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    if leaf in ("weight", "weight_orig"):
        if len(shape) == 4:
            transposed = ".upconv_block" in key or ".synthesis_" in key
            fan_in = (shape[0] if transposed else shape[1]) * shape[2] * shape[3]
            stride2_transposed = transposed and not (".synthesis_" in key and ".conv3." in key)
            if stride2_transposed:
                fan_in /= 4.0  # each output pixel sees ~k*k/4 taps of a stride-2 transposed conv
            return torch.randn(shape, generator=g) * math.sqrt(1.5 / fan_in)
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "bias":
        if key == "Hyperprior.synthesis_std.conv3.bias":
            return 0.4 + 0.05 * torch.randn(shape, generator=g)   # keep part of the scales above MIN_SCALE
        return torch.randn(shape, generator=g) * 0.05
    if leaf == "gamma":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "beta":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf in ("weight_u", "weight_v"):
        v = torch.randn(shape, generator=g)
        return v / v.norm()
    if leaf.startswith("H_"):
        k = int(leaf[2:])
        filt = (1, 3, 3, 3, 1)
        scale = 10.0 ** (1 / 4)
        init = math.log(math.expm1(1 / scale / filt[k + 1]))       # hyperprior_model.py:289
        return init + 0.2 * torch.randn(shape, generator=g)
    if leaf.startswith("a_"):
        return 0.3 * torch.randn(shape, generator=g)
    if leaf.startswith("b_"):
        return torch.rand(shape, generator=g) - 0.5
    raise KeyError(key)


def synth_state_dict(seed=0, **kw):
    sd = {k: synth_tensor(k, shp, seed) for k, shp in hific_shapes(**kw).items()}
    # spectral-norm buffers of a trained discriminator are (nearly) converged singular vectors; random unit vectors
    # would make sigma = u.W v tiny and W / sigma explode.  Converge them with a few power iterations.
    for k in [k for k in sd if k.endswith(".weight_u")]:
        w = sd[k.replace("weight_u", "weight_orig")]
        wm = w.reshape(w.shape[0], -1)
        u, v = sd[k], sd[k.replace("weight_u", "weight_v")]
        for _ in range(12):
            v = torch.nn.functional.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12)
            u = torch.nn.functional.normalize(torch.mv(wm, v), dim=0, eps=1e-12)
        sd[k], sd[k.replace("weight_u", "weight_v")] = u, v
    return sd


def synth_image(n, h, w, seed=0, channels=3):
    """Smooth-ish image in [0, 1]: low-frequency pattern plus noise (so the latents are not pure noise)."""
    g = _gen(f"image{n}x{h}x{w}", seed)
    yy = torch.linspace(0, 1, h).view(1, 1, h, 1)
    xx = torch.linspace(0, 1, w).view(1, 1, 1, w)
    ph = torch.rand(n, channels, 1, 1, generator=g) * 6.28
    base = 0.5 + 0.25 * torch.sin(6.28 * (2 * xx + 3 * yy) + ph) + 0.15 * torch.cos(6.28 * (5 * xx - 4 * yy) + 2 * ph)
    noise = 0.1 * torch.randn(n, channels, h, w, generator=g)
    return (base + noise).clamp(0, 1).contiguous()


def synth_noise(shape, tag, seed=0):
    """U(-1/2, 1/2) quantisation noise (the reference draws it with torch.nn.init.uniform_, hyperprior.py:65)."""
    return torch.rand(shape, generator=_gen("noise" + tag, seed)) - 0.5


def instance_norm_variant(sd):
    """The same synthetic parameters for the use_channel_norm = False architecture: torch.nn.InstanceNorm2d keeps its
    affine pair as `weight` / `bias` of shape (c,) where ChannelNorm2D has `gamma` / `beta` of shape (1, c, 1, 1)
    (src/normalisation/instance.py:7-15 vs channel.py:29-46)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".gamma"):
            out[k[:-len(".gamma")] + ".weight"] = v.reshape(-1).clone()
        elif k.endswith(".beta"):
            out[k[:-len(".beta")] + ".bias"] = v.reshape(-1).clone()
        else:
            out[k] = v
    return out
