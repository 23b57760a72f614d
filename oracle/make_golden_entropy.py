"""Generate tests/golden/entropy_*.npz by running the REAL reference's entropy-coding modules (imported from
/root/reference through oracle/ref_shim.install_ans) -- TEST INFRASTRUCTURE, build container only.

    python oracle/make_golden_entropy.py

Fixtures: the prior tables (gaussian + logistic), the hyperprior tables / tails for the synthetic density
parameters, bitstreams + decoded symbols of the vectorised coder for batch 1 and batch 2 (with out-of-range
symbols so the overflow path is exercised), scale indices, Shannon bit estimates, and one .hfc container.
The script ends by checking oracle/entropy_oracle.py against everything it wrote (bit-exact).
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def synth_latents(shape, seed, outliers=True):
    """(y, mean, scale) with a few symbols far outside the table range of their scale."""
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(shape, generator=g) * 3.0
    mu = torch.randn(shape, generator=g)
    sc = (torch.rand(shape, generator=g) * 2.0) ** 2 + 0.02       # crosses the 0.11 bound and several table rows
    if outliers:
        flat = y.view(-1)
        pos = torch.randperm(flat.numel(), generator=g)[:max(4, flat.numel() // 40)]
        flat[pos] += torch.randn(pos.numel(), generator=g) * 25.0
    return y, mu, sc


def main():
    ref_shim.install_ans()
    from src.compression import compression_utils, hyperprior_model, prior_model
    from src.helpers import maths
    from hific_b200 import synth
    from oracle import entropy_oracle as EO
    quiet = contextlib.redirect_stdout(io.StringIO())
    out = {}

    # ---------------------------------------------------------------- prior tables
    models = {}
    for kind in ("gaussian", "logistic"):
        density = prior_model.PriorDensity(n_channels=8, likelihood_type=kind)
        with quiet:
            pem = prior_model.PriorEntropyModel(distribution=density)
        models[kind] = pem
        out[f"prior_{kind}.CDF"] = pem.CDF.numpy().astype(np.int32)
        out[f"prior_{kind}.CDF_offset"] = pem.CDF_offset.numpy()
        out[f"prior_{kind}.CDF_length"] = pem.CDF_length.numpy()
        out[f"prior_{kind}.scale_table"] = pem.scale_table_tensor.numpy()

    # ---------------------------------------------------------------- prior model: indices, bits, bitstreams
    pem = models["gaussian"]
    for name, shape in (("b1", (1, 8, 5, 7)), ("b2", (2, 8, 4, 6)), ("b3", (3, 8, 3, 3))):
        y, mu, sc = synth_latents(shape, seed=len(name) + shape[0])
        sc_b = maths.LowerBoundToward.apply(sc, 0.11)
        with quiet:
            enc, coding_shape, rounded = pem.compress(y, mu, sc_b, vectorize=True, block_encode=True)
            dec, raw = pem.decompress(enc, mu, sc_b, broadcast_shape=shape[2:], coding_shape=coding_shape,
                                      vectorize=True, block_decode=True)
        bits, bpp, bpi = pem._estimate_compression_bits(y, mu, sc_b, spatial_shape=(64, 64))
        out[f"prior_{name}.y"], out[f"prior_{name}.mean"], out[f"prior_{name}.scale"] = y.numpy(), mu.numpy(), sc.numpy()
        out[f"prior_{name}.indices"] = pem.compute_indices(sc_b).numpy()
        out[f"prior_{name}.symbols"] = rounded.numpy()
        out[f"prior_{name}.encoded"] = np.asarray(enc, dtype=np.uint32)
        out[f"prior_{name}.coding_shape"] = np.array(coding_shape, dtype=np.int64)
        out[f"prior_{name}.decoded_raw"] = raw.numpy()
        out[f"prior_{name}.decoded"] = dec.numpy()
        out[f"prior_{name}.bits"] = np.array(float(bits))
        print(name, "words", len(enc), "round-trip exact fraction", float((raw == rounded).float().mean()))

    # ---------------------------------------------------------------- hyperprior tables (synthetic density, seed 0)
    sd = synth.synth_state_dict(0)
    c = 320
    density = hyperprior_model.HyperpriorDensity(n_channels=c)
    dparams = {k.split(".")[-1]: v for k, v in sd.items() if k.startswith("Hyperprior.hyperlatent_likelihood.")}
    density.load_state_dict(dparams, strict=True)
    # the reference evaluates the tails on utils.get_device(): CPU in this container
    hem = hyperprior_model.HyperpriorEntropyModel(distribution=density)
    with quiet:
        hem.build_tables()
    out["hyper.CDF"] = hem.CDF.numpy().astype(np.int32)
    out["hyper.CDF_offset"] = hem.CDF_offset.numpy()
    out["hyper.CDF_length"] = hem.CDF_length.numpy()
    out["hyper.median"] = hem.medians.reshape(-1).numpy()
    out["hyper.lower_tail"] = density.lower_tail(hem.tail_mass).cpu().numpy()
    out["hyper.upper_tail"] = density.upper_tail(hem.tail_mass).cpu().numpy()
    for name, shape in (("b1", (1, c, 3, 2)), ("b2", (2, c, 2, 2))):
        g = torch.Generator().manual_seed(7 + shape[0])
        z = torch.randn(shape, generator=g) * 4.0
        z.view(-1)[::37] += 40.0                                   # overflow symbols
        z.view(-1)[5::53] -= 33.0
        with quiet:
            enc, coding_shape, rounded = hem.compress(z, vectorize=True, block_encode=True)
            dec, raw = hem.decompress(enc, batch_shape=shape[0], broadcast_shape=shape[2:], coding_shape=coding_shape,
                                      vectorize=True, block_decode=True)
        bits, bpp, bpi = hem._estimate_compression_bits(z, spatial_shape=(64, 64))
        out[f"hyper_{name}.z"] = z.numpy()
        out[f"hyper_{name}.symbols"] = rounded.numpy()
        out[f"hyper_{name}.encoded"] = np.asarray(enc, dtype=np.uint32)
        out[f"hyper_{name}.coding_shape"] = np.array(coding_shape, dtype=np.int64)
        out[f"hyper_{name}.decoded_raw"] = raw.numpy()
        out[f"hyper_{name}.bits"] = np.array(float(bits.detach()))
        print("hyper", name, "words", len(enc), "round-trip exact fraction", float((raw == rounded).float().mean()))

    # ---------------------------------------------------------------- .hfc container
    co = compression_utils.CompressionOutput(
        hyperlatents_encoded=out["hyper_b1.encoded"], latents_encoded=out["prior_b1.encoded"],
        hyperlatent_spatial_shape=(3, 2), batch_shape=1, spatial_shape=(80, 112),
        hyper_coding_shape=tuple(out["hyper_b1.coding_shape"]), latent_coding_shape=tuple(out["prior_b1.coding_shape"]))
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "x.hfc")
        try:
            compression_utils.save_compressed_format(co, p)
        except AttributeError:
            pass        # the reference then reads compression_output.total_bpp, which this namedtuple lacks: file is complete
        out["container.bytes"] = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
        back = compression_utils.load_compressed_format(p)
        assert np.array_equal(back.latents_encoded, out["prior_b1.encoded"])

    # ---------------------------------------------------------------- whole Model.compress / decompress (CPU reference)
    import logging
    from default_config import ModelModes, ModelTypes, mse_lpips_args
    from src.model import Model

    class A(mse_lpips_args):
        pass
    cfg = A()
    cfg.image_dims, cfg.latent_dims, cfg.batch_size = (3, 256, 256), (cfg.latent_channels, 16, 16), 1
    with quiet:
        model = Model(cfg, logging.getLogger("golden"), model_mode=ModelModes.EVALUATION, model_type=ModelTypes.COMPRESSION)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    model.eval()
    with quiet:
        model.Hyperprior.hyperprior_entropy_model.build_tables()
    assert np.array_equal(model.Hyperprior.hyperprior_entropy_model.CDF.numpy(), out["hyper.CDF"])
    for name, (b, h, w) in (("m1", (1, 100, 144)), ("m2", (2, 96, 128))):
        x = synth.synth_image(b, h, w, 30 + b)
        with torch.no_grad(), quiet:
            co = model.compress(x, silent=True)
            rec = model.decompress(co)
        out[f"model_{name}.hyperlatents_encoded"] = np.asarray(co.hyperlatents_encoded, dtype=np.uint32)
        out[f"model_{name}.latents_encoded"] = np.asarray(co.latents_encoded, dtype=np.uint32)
        out[f"model_{name}.shapes"] = np.array(list(co.hyperlatent_spatial_shape) + list(co.spatial_shape)
                                               + list(co.hyper_coding_shape) + list(co.latent_coding_shape)
                                               + [co.batch_shape], dtype=np.int64)
        out[f"model_{name}.bpp"] = np.array([co.hyperlatent_bpp, co.latent_bpp, co.total_bpp, co.hyperlatent_bits,
                                             co.latent_bits, co.total_bits], dtype=np.float64)
        out[f"model_{name}.reconstruction"] = rec.numpy()
        print("model", name, "words", len(co.hyperlatents_encoded), len(co.latents_encoded), "bpp", co.total_bpp)

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, "entropy_coding.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

    # ---------------------------------------------------------------- oracle vs reference, bit-exact
    g = np.load(path)
    for kind in ("gaussian", "logistic"):
        cdf, off, length, table = EO.prior_tables(kind)
        assert np.array_equal(cdf, g[f"prior_{kind}.CDF"]) and np.array_equal(off, g[f"prior_{kind}.CDF_offset"])
        assert np.array_equal(length, g[f"prior_{kind}.CDF_length"]) and np.array_equal(table.numpy(), g[f"prior_{kind}.scale_table"])
    tables = EO.prior_tables("gaussian")
    for name in ("b1", "b2", "b3"):
        y, mu, sc = (torch.from_numpy(g[f"prior_{name}.{k}"]) for k in ("y", "mean", "scale"))
        sc_b = torch.clamp(sc, 0.11)
        enc, cs, sym = EO.prior_compress(y, mu, sc_b, tables)
        assert np.array_equal(enc, g[f"prior_{name}.encoded"]), name
        assert tuple(cs) == tuple(g[f"prior_{name}.coding_shape"])
        dec, raw = EO.prior_decompress(enc, mu, sc_b, tables)
        assert np.array_equal(raw.numpy(), g[f"prior_{name}.decoded_raw"]) and np.array_equal(dec.numpy(), g[f"prior_{name}.decoded"])
        assert float(EO.prior_bits(y, mu, sc_b)) == float(g[f"prior_{name}.bits"])
    ht = EO.hyper_tables(dparams)
    for k, v in zip(("CDF", "CDF_offset", "CDF_length", "lower_tail", "upper_tail", "median"), ht):
        assert np.array_equal(v, g[f"hyper.{k}"]), k
    for name in ("b1", "b2"):
        z = torch.from_numpy(g[f"hyper_{name}.z"])
        enc, cs, sym = EO.hyper_compress(z, ht)
        assert np.array_equal(enc, g[f"hyper_{name}.encoded"]), name
        assert np.array_equal(EO.hyper_decompress(enc, z.shape, ht).numpy(), g[f"hyper_{name}.decoded_raw"])
        assert float(EO.hyper_bits(z, dparams)) == float(g[f"hyper_{name}.bits"])
    cb = EO.container_bytes((3, 2), (80, 112), tuple(g["hyper_b1.coding_shape"]), tuple(g["prior_b1.coding_shape"]), 1,
                            g["hyper_b1.encoded"], g["prior_b1.encoded"])
    assert cb == g["container.bytes"].tobytes()
    print("oracle == reference on every entropy-coding fixture (bit-exact)")


if __name__ == "__main__":
    main()
