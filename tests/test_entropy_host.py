"""CPU: the entropy-coding side of the compress path.

  * the oracle (oracle/entropy_oracle.py) against the golden vectors produced by the REAL reference's coder
    (tests/golden/entropy_coding.npz, oracle/make_golden_entropy.py) -- this is what pins the oracle;
  * the product's host code (csrc/entropy_host.cpp through hific_b200.compression) against the same golden vectors
    and, on random inputs, against the oracle: tables, bitstreams, decoded symbols and the .hfc container, all
    bit-exact.  The host coder is product code that runs on the CPU by design (BASELINE north_star: "the sequential
    ANS entropy coder stays on the host"); nothing here touches a GPU.
"""
import os

import numpy as np
import pytest
import torch

from hific_b200 import synth
from hific_b200._lib import HfcError
from hific_b200.compression import compression_utils, entropy_coding, hyperprior_model, prior_model
from oracle import entropy_oracle as EO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "entropy_coding.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def density_params():
    sd = synth.synth_state_dict(0)
    return {k.split(".")[-1]: v for k, v in sd.items() if k.startswith("Hyperprior.hyperlatent_likelihood.")}


def prior_tables_of(gold, kind="gaussian"):
    return entropy_coding.Tables(gold[f"prior_{kind}.CDF"], gold[f"prior_{kind}.CDF_length"], gold[f"prior_{kind}.CDF_offset"])


# ------------------------------------------------------------------------------------------------ oracle vs reference
def test_oracle_prior_tables_match_reference(gold):
    cdf, off, length, table = EO.prior_tables("gaussian")
    assert np.array_equal(cdf, gold["prior_gaussian.CDF"])
    assert np.array_equal(off, gold["prior_gaussian.CDF_offset"])
    assert np.array_equal(length, gold["prior_gaussian.CDF_length"])
    assert np.array_equal(table.numpy(), gold["prior_gaussian.scale_table"])


@pytest.mark.parametrize("name", ["b1", "b2", "b3"])
def test_oracle_prior_coder_matches_reference(gold, name):
    tables = (gold["prior_gaussian.CDF"], gold["prior_gaussian.CDF_offset"], gold["prior_gaussian.CDF_length"],
              torch.from_numpy(gold["prior_gaussian.scale_table"]))
    y, mu, sc = (torch.from_numpy(gold[f"prior_{name}.{k}"]) for k in ("y", "mean", "scale"))
    sc = torch.clamp(sc, 0.11)
    assert np.array_equal(EO.compute_indices(sc, tables[3]).numpy(), gold[f"prior_{name}.indices"])
    enc, coding_shape, sym = EO.prior_compress(y, mu, sc, tables)
    assert np.array_equal(sym, gold[f"prior_{name}.symbols"])
    assert np.array_equal(enc, gold[f"prior_{name}.encoded"])
    assert tuple(coding_shape) == tuple(gold[f"prior_{name}.coding_shape"])
    dec, raw = EO.prior_decompress(enc, mu, sc, tables)
    assert np.array_equal(raw.numpy(), gold[f"prior_{name}.decoded_raw"])
    assert np.array_equal(dec.numpy(), gold[f"prior_{name}.decoded"])
    assert float(EO.prior_bits(y, mu, sc)) == float(gold[f"prior_{name}.bits"])
    # the fixtures exercise the escape path, including the reference's lossy wide escapes
    assert (raw.numpy() != gold[f"prior_{name}.symbols"]).any()


def test_oracle_hyper_tables_and_coder_match_reference(gold, density_params):
    ht = EO.hyper_tables(density_params)
    for k, v in zip(("CDF", "CDF_offset", "CDF_length", "lower_tail", "upper_tail", "median"), ht):
        assert np.array_equal(v, gold[f"hyper.{k}"]), k
    for name in ("b1", "b2"):
        z = torch.from_numpy(gold[f"hyper_{name}.z"])
        enc, coding_shape, sym = EO.hyper_compress(z, ht)
        assert np.array_equal(enc, gold[f"hyper_{name}.encoded"])
        assert np.array_equal(EO.hyper_decompress(enc, z.shape, ht).numpy(), gold[f"hyper_{name}.decoded_raw"])
        assert float(EO.hyper_bits(z, density_params)) == float(gold[f"hyper_{name}.bits"])


def test_oracle_container_matches_reference(gold):
    cb = EO.container_bytes((3, 2), (80, 112), tuple(gold["hyper_b1.coding_shape"]), tuple(gold["prior_b1.coding_shape"]),
                            1, gold["hyper_b1.encoded"], gold["prior_b1.encoded"])
    assert cb == gold["container.bytes"].tobytes()


# ------------------------------------------------------------------------------------------------ product host code
@pytest.mark.parametrize("kind", ["gaussian", "logistic"])
def test_product_prior_tables_match_reference(gold, kind):
    pem = prior_model.PriorEntropyModel(distribution=prior_model.PriorDensity(220, likelihood_type=kind))
    assert np.array_equal(pem.CDF.numpy(), gold[f"prior_{kind}.CDF"])
    assert np.array_equal(pem.CDF_offset.numpy(), gold[f"prior_{kind}.CDF_offset"])
    assert np.array_equal(pem.CDF_length.numpy(), gold[f"prior_{kind}.CDF_length"])
    assert np.array_equal(pem.scale_table_tensor.numpy(), gold[f"prior_{kind}.scale_table"])
    assert pem.CDF.dtype == torch.int32 and not pem.CDF.requires_grad
    assert set(pem.state_dict()) == {"CDF", "CDF_offset", "CDF_length", "scale_table_tensor", "min_scale_tensor"}


def test_product_hyper_tables_match_reference(gold, density_params):
    d = hyperprior_model.HyperpriorDensity(320)
    d.load_state_dict(density_params)
    hem = hyperprior_model.HyperpriorEntropyModel(d)
    assert "CDF" not in hem.state_dict()                      # as in the reference: registered by build_tables()
    hem.build_tables()
    assert np.array_equal(hem.CDF.numpy(), gold["hyper.CDF"])
    assert np.array_equal(hem.CDF_offset.numpy(), gold["hyper.CDF_offset"])
    assert np.array_equal(hem.CDF_length.numpy(), gold["hyper.CDF_length"])
    assert np.array_equal(hem.medians.reshape(-1).numpy(), gold["hyper.median"])
    assert "distribution.H_0" in hem.state_dict() and "CDF" in hem.state_dict()


def test_pmf_to_quantized_cdf_matches_oracle_on_random_rows():
    g = torch.Generator().manual_seed(5)
    for trial in range(40):
        n = int(torch.randint(2, 400, (1,), generator=g))
        pmf = torch.rand(n, generator=g) ** 6                  # many tiny entries -> zero frequencies to repair
        if trial % 3 == 0:
            pmf[torch.rand(n, generator=g) < 0.3] = 0.0
        pmf[int(torch.randint(0, n, (1,), generator=g))] += 0.5
        pmf = pmf / pmf.sum() * (0.9 if trial % 2 else 1.0)
        want = EO.pmf_to_quantized_cdf(pmf, 16)
        got = entropy_coding.pmf_to_quantized_cdf(pmf.numpy(), 16)
        assert np.array_equal(got, want), trial
        assert got[0] == 0 and got[-1] == 65536 and (np.diff(got) > 0).all()
    with pytest.raises(HfcError):
        entropy_coding.pmf_to_quantized_cdf(np.array([0.5, -0.1, 0.6], dtype=np.float32))
    with pytest.raises(HfcError):
        entropy_coding.pmf_to_quantized_cdf(np.zeros(4, dtype=np.float32))


@pytest.mark.parametrize("name", ["b1", "b2", "b3"])
def test_host_coder_reproduces_reference_prior_messages(gold, name):
    T = prior_tables_of(gold)
    sym, idx = gold[f"prior_{name}.symbols"], gold[f"prior_{name}.indices"]
    steps, lanes, coding_shape = prior_model.coder_shape(sym.shape)
    s, _ = EO.to_coder_layout(sym)
    i, _ = EO.to_coder_layout(idx)
    assert s.shape == (steps, lanes) and tuple(coding_shape) == tuple(gold[f"prior_{name}.coding_shape"])
    enc = entropy_coding.vec_ans_index_encoder(s, i, T)
    assert enc.dtype == np.uint32 and np.array_equal(enc, gold[f"prior_{name}.encoded"])
    dec = entropy_coding.vec_ans_index_decoder(gold[f"prior_{name}.encoded"], i, T)
    assert np.array_equal(prior_model.coder_to_nchw(dec, sym.shape), gold[f"prior_{name}.decoded_raw"].astype(np.int32))


@pytest.mark.parametrize("name", ["b1", "b2"])
def test_host_coder_reproduces_reference_hyper_messages(gold, name):
    T = entropy_coding.Tables(gold["hyper.CDF"], gold["hyper.CDF_length"], gold["hyper.CDF_offset"])
    sym = gold[f"hyper_{name}.symbols"]
    s, _ = EO.to_coder_layout(sym)
    i, _ = EO.to_coder_layout(EO.hyper_indices(sym.shape))
    hem = hyperprior_model.HyperpriorEntropyModel(hyperprior_model.HyperpriorDensity(320))
    assert np.array_equal(hem._coder_indices(sym.shape), i)
    enc = entropy_coding.vec_ans_index_encoder(s, i, T)
    assert np.array_equal(enc, gold[f"hyper_{name}.encoded"])
    dec = entropy_coding.vec_ans_index_decoder(enc, i, T)
    assert np.array_equal(prior_model.coder_to_nchw(dec, sym.shape), gold[f"hyper_{name}.decoded_raw"].astype(np.int32))


@pytest.mark.parametrize("steps,lanes,seed", [(1, 1, 0), (7, 1, 1), (1, 300, 2), (33, 220, 3), (2, 4096, 4)])
def test_host_coder_matches_oracle_on_random_symbols(gold, steps, lanes, seed):
    """Ragged shapes, every table row, in-range symbols (must round-trip exactly) and escapes (must follow the
    reference's lowest-nibble quirk -- compared with the oracle, which is pinned to the reference)."""
    T = prior_tables_of(gold)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 64, size=(steps, lanes)).astype(np.int32)
    half = (T.length[idx] - 2) // 2
    sym = np.rint(rng.normal(size=(steps, lanes)) * np.maximum(1, half / 3)).astype(np.int32)
    far = rng.random((steps, lanes)) < 0.03
    sym[far] += rng.integers(-4000, 4000, size=int(far.sum())).astype(np.int32)
    enc = entropy_coding.vec_ans_index_encoder(sym, idx, T)
    want = EO.vec_encode(sym, idx, T.cdf, T.length, T.offset)
    assert np.array_equal(enc, want)
    dec = entropy_coding.vec_ans_index_decoder(enc, idx, T)
    assert np.array_equal(dec, EO.vec_decode(enc, idx, T.cdf, T.length, T.offset))
    inside = (sym - T.offset[idx] >= 0) & (sym - T.offset[idx] < T.length[idx] - 2)
    assert np.array_equal(dec[inside], sym[inside])
    assert len(enc) >= 2 * lanes


def test_host_coder_rejects_bad_input(gold):
    T = prior_tables_of(gold)
    sym = np.zeros((2, 8), dtype=np.int32)
    idx = np.zeros((2, 8), dtype=np.int32)
    enc = entropy_coding.vec_ans_index_encoder(sym, idx, T)
    with pytest.raises(HfcError):
        entropy_coding.vec_ans_index_encoder(sym, idx + 64, T)                    # table row out of range
    with pytest.raises(HfcError):
        entropy_coding.vec_ans_index_decoder(enc[:10], idx, T)                    # shorter than the lane states
    bad = entropy_coding.Tables(T.cdf, T.length + 5000, T.offset)                 # rows longer than the table
    with pytest.raises(HfcError):
        entropy_coding.vec_ans_index_encoder(sym, idx, bad)


def test_container_round_trip_and_wire_format(gold, tmp_path):
    co = compression_utils.CompressionOutput(
        hyperlatents_encoded=gold["hyper_b1.encoded"], latents_encoded=gold["prior_b1.encoded"],
        hyperlatent_spatial_shape=(3, 2), batch_shape=1, spatial_shape=(80, 112),
        hyper_coding_shape=tuple(gold["hyper_b1.coding_shape"]), latent_coding_shape=tuple(gold["prior_b1.coding_shape"]),
        total_bpp=0.25)
    p = str(tmp_path / "x.hfc")
    actual_bpp, theoretical_bpp = compression_utils.save_compressed_format(co, p)
    data = open(p, "rb").read()
    assert data == gold["container.bytes"].tobytes()                              # byte-identical to the reference's file
    assert theoretical_bpp == 0.25 and abs(actual_bpp - 8 * len(data) / (80 * 112)) < 1e-12
    back = compression_utils.load_compressed_format(p)
    assert np.array_equal(back.hyperlatents_encoded, co.hyperlatents_encoded)
    assert np.array_equal(back.latents_encoded, co.latents_encoded)
    assert back.hyperlatent_spatial_shape == (3, 2) and back.spatial_shape == (80, 112) and back.batch_shape == 1
    assert back.hyper_coding_shape == co.hyper_coding_shape and back.latent_coding_shape == co.latent_coding_shape


def test_evaluation_model_has_the_reference_state_dict_keys():
    """EVALUATION mode adds Hyperprior.hyperprior_entropy_model.distribution.* (aliases of the density
    parameters) and Hyperprior.prior_entropy_model.{CDF, CDF_offset, CDF_length, scale_table_tensor, min_scale_tensor}."""
    import logging
    from hific_b200.config import ModelModes, mse_lpips_args
    from hific_b200.model import Model
    m = Model(mse_lpips_args(), logging.getLogger("t"), model_mode=ModelModes.EVALUATION)
    keys = set(m.state_dict())
    train_keys = set(synth.synth_state_dict(0))
    extra = keys - train_keys
    want = {f"Hyperprior.hyperprior_entropy_model.distribution.{p}_{k}" for p in "Hab" for k in range(4)}
    want |= {"Hyperprior.prior_entropy_model." + k for k in ("CDF", "CDF_offset", "CDF_length", "scale_table_tensor",
                                                           "min_scale_tensor")}
    assert extra == want and train_keys <= keys
    sd = m.state_dict()
    assert sd["Hyperprior.hyperprior_entropy_model.distribution.H_0"].data_ptr() == \
        sd["Hyperprior.hyperlatent_likelihood.H_0"].data_ptr()
    assert tuple(sd["Hyperprior.prior_entropy_model.CDF"].shape) == (64, 1481)


def _random_tables(rng, rows, max_len, precision=16):
    """Valid quantised CDF rows of ragged lengths built with the product's own quantiser from random PMFs."""
    lengths = rng.integers(1, max_len + 1, size=rows)                 # pmf_length >= 1 -> cdf_length = pmf_length + 2 >= 3
    cdf = np.zeros((rows, max_len + 2), dtype=np.int32)
    for r in range(rows):
        pmf = rng.random(lengths[r] + 1).astype(np.float32) ** 4 + 1e-7   # + the overflow / tail bin
        pmf /= pmf.sum()
        cdf[r, :lengths[r] + 2] = entropy_coding.pmf_to_quantized_cdf(pmf, precision)
    offset = -rng.integers(0, max_len, size=rows).astype(np.int32)
    return entropy_coding.Tables(cdf, (lengths + 2).astype(np.int32), offset)


@pytest.mark.parametrize("seed", range(12))
def test_host_coder_property_random_tables(seed):
    """Random ragged tables (1..40 symbols per row, including single-symbol rows), random shapes, symbols inside and far
    outside every row: the C coder equals the oracle word for word, decodes what the oracle decodes, and every in-table
    symbol round-trips exactly."""
    rng = np.random.default_rng(100 + seed)
    rows = int(rng.integers(1, 9))
    T = _random_tables(rng, rows, int(rng.integers(1, 41)))
    steps, lanes = int(rng.integers(1, 20)), int(rng.integers(1, 70))
    idx = rng.integers(0, rows, size=(steps, lanes)).astype(np.int32)
    span = (T.length[idx] - 2).astype(np.int64)
    sym = (T.offset[idx] + rng.integers(0, np.maximum(span, 1))).astype(np.int32)
    wild = rng.random((steps, lanes)) < 0.15
    sym[wild] += rng.integers(-300, 300, size=int(wild.sum())).astype(np.int32)
    enc = entropy_coding.vec_ans_index_encoder(sym, idx, T)
    assert np.array_equal(enc, EO.vec_encode(sym, idx, T.cdf, T.length, T.offset))
    dec = entropy_coding.vec_ans_index_decoder(enc, idx, T)
    assert np.array_equal(dec, EO.vec_decode(enc, idx, T.cdf, T.length, T.offset))
    inside = (sym - T.offset[idx] >= 0) & (sym - T.offset[idx] < span)
    assert np.array_equal(dec[inside], sym[inside])
