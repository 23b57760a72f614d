"""Compress-path helpers with the reference's names (src/compression/compression_utils.py): the
`CompressionOutput` record, tail estimation for the factorized density, and the `.hfc` container
(wire-compatible with the reference: files written by either side load on the other).
"""
import os
from collections import namedtuple

import numpy as np
import torch

# Random bits for fencing in the bitstream (compression_utils.py:17)
_MAGIC_VALUE_SEP = b'\x46\xE2\x84\x92'

# The reference declares two records of this name: the full one returned by Hyperprior.compress_forward
# (src/hyperprior.py:25-40) and a 7-field one used by load_compressed_format (compression_utils.py:21-29).
CompressionOutput = namedtuple(
    "CompressionOutput",
    ["hyperlatents_encoded", "latents_encoded", "hyperlatent_spatial_shape", "batch_shape", "spatial_shape",
     "hyper_coding_shape", "latent_coding_shape", "hyperlatent_bits", "latent_bits", "total_bits",
     "hyperlatent_bpp", "latent_bpp", "total_bpp"],
    defaults=[None] * 6)


def estimate_tails(cdf, target, shape, dtype=torch.float32, extra_counts=24):
    """compression_utils.py:35-85: Adam iteration on |cdf(x) - target| from x = 0, stopped `extra_counts` steps after
    every element has crossed its optimum.  Runs on the CPU (init-time table building; the reference uses
    utils.get_device()): the tables must be reproducible on the decoding machine."""
    lr, eps = 1e-2, 1e-8
    beta_1, beta_2 = 0.9, 0.99
    tails = torch.zeros(shape, dtype=dtype, requires_grad=True)
    m = torch.zeros(shape, dtype=dtype)
    v = torch.ones(shape, dtype=dtype)
    counts = torch.zeros(shape, dtype=torch.int32)
    while torch.min(counts) < extra_counts:
        loss = abs(cdf(tails) - target)
        loss.backward(torch.ones_like(tails))
        tgrad = tails.grad
        with torch.no_grad():
            m = beta_1 * m + (1. - beta_1) * tgrad
            v = beta_2 * v + (1. - beta_2) * torch.square(tgrad)
            tails -= lr * m / (torch.sqrt(v) + eps)
        counts = torch.where(torch.logical_or(counts > 0, tgrad * tails > 0), counts + 1, counts)
        tails.grad.zero_()
    return tails


def check_argument_shapes(cdf, cdf_length, cdf_offset):
    """compression_utils.py:109-120."""
    if len(cdf.size()) != 2 or cdf.size(1) < 3:
        raise ValueError("'cdf' should be 2-D and cdf.dim_size(1) >= 3: ", cdf.size())
    if len(cdf_length.size()) != 1 or cdf_length.size(0) != cdf.size(0):
        raise ValueError("'cdf_length' should be 1-D and its length should match the number of rows in 'cdf': ",
                         cdf_length.size())
    if len(cdf_offset.size()) != 1 or cdf_offset.size(0) != cdf.size(0):
        raise ValueError("'cdf_offset' should be 1-D and its length should match the number of rows in 'cdf': ",
                         cdf_offset.size())


def _write_shapes(shape, f):
    for s in shape:
        assert 0 <= int(s) < 2 ** 16, s
        f.write(np.uint16(s).tobytes())


def _read_shapes(f, n):
    return tuple(int(v) for v in np.frombuffer(f.read(2 * n), np.uint16, count=n))


def _write_message(msg, f):
    msg = np.ascontiguousarray(msg, dtype=np.uint32)
    assert msg.nbytes < 2 ** 32
    f.write(np.uint32(msg.nbytes).tobytes())
    f.write(msg.tobytes())
    f.write(_MAGIC_VALUE_SEP)


def _read_message(f):
    nbytes = int(np.frombuffer(f.read(4), np.uint32, count=1)[0])
    msg = np.frombuffer(f.read(nbytes), np.uint32, count=-1)
    assert f.read(4) == _MAGIC_VALUE_SEP, "corrupt .hfc file"
    return msg


def save_compressed_format(compression_output, out_path):
    """compression_utils.py:300-335: shapes as uint16, magic separator, the two uint32 messages.  Returns
    (actual_bpp, theoretical_bpp)."""
    co = compression_output
    with open(out_path, 'wb') as f:
        _write_shapes(co.hyperlatent_spatial_shape, f)
        _write_shapes(co.spatial_shape, f)
        _write_shapes(co.hyper_coding_shape, f)
        _write_shapes(co.latent_coding_shape, f)
        _write_shapes([co.batch_shape], f)
        f.write(_MAGIC_VALUE_SEP)
        _write_message(co.hyperlatents_encoded, f)
        _write_message(co.latents_encoded, f)
    actual_bpp = 8. * float(os.path.getsize(out_path)) / np.prod(co.spatial_shape)
    tb = co.total_bpp
    theoretical_bpp = float('nan') if tb is None else float(tb.item() if hasattr(tb, "item") else tb)
    return actual_bpp, theoretical_bpp


def load_compressed_format(in_path):
    """compression_utils.py:337-371."""
    with open(in_path, 'rb') as f:
        hyperlatent_spatial_shape = _read_shapes(f, 2)
        spatial_shape = _read_shapes(f, 2)
        hyper_coding_shape = _read_shapes(f, 3)
        latent_coding_shape = _read_shapes(f, 3)
        batch_shape = _read_shapes(f, 1)
        assert f.read(4) == _MAGIC_VALUE_SEP, "not a .hfc file"
        hyperlatents_encoded = _read_message(f)
        latents_encoded = _read_message(f)
    return CompressionOutput(hyperlatents_encoded=hyperlatents_encoded, latents_encoded=latents_encoded,
                             hyperlatent_spatial_shape=hyperlatent_spatial_shape, spatial_shape=spatial_shape,
                             hyper_coding_shape=hyper_coding_shape, latent_coding_shape=latent_coding_shape,
                             batch_shape=batch_shape[0])
