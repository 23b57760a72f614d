"""Host side of the entropy coder: numpy views over the C entry points of libhfc (csrc/entropy_host.cpp), which
replace the reference's Python coder (src/compression/entropy_coding.py:251-476, 555-676; src/compression/ans.py)
and its `maths.pmf_to_quantized_cdf` (src/helpers/maths.py:5-73).  Bit-compatible with the reference: messages
produced by either side decode on the other.  No GPU involved ("the sequential ANS entropy coder stays on the host").
"""
import ctypes

import numpy as np

from .._lib import check, lib

OVERFLOW_WIDTH = 4
PATCH_SIZE = (1, 1)
PRECISION_P = 16


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def pmf_to_quantized_cdf(pmf, precision=PRECISION_P):
    """(n,) float32 probabilities -> (n + 1,) int32 quantised CDF summing to 2**precision."""
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    out = np.empty(pmf.shape[0] + 1, dtype=np.int32)
    check(lib.hfc_pmf_to_quantized_cdf_host(_p(pmf), pmf.shape[0], int(precision), _p(out)), "pmf_to_quantized_cdf")
    return out


class Tables:
    """Host copy of (CDF, CDF_length, CDF_offset) in the form the coder wants."""

    def __init__(self, cdf, cdf_length, cdf_offset):
        self.cdf, self.length, self.offset = _i32(cdf), _i32(cdf_length), _i32(cdf_offset)
        assert self.cdf.ndim == 2 and self.length.shape == self.offset.shape == (self.cdf.shape[0],)


def vec_ans_index_encoder(symbols, indices, tables, precision=PRECISION_P):
    """symbols / indices: int32 [steps][lanes] in coder order -> flat uint32 message."""
    symbols, indices = _i32(symbols), _i32(indices)
    assert symbols.shape == indices.shape and symbols.ndim == 2
    steps, lanes = symbols.shape
    args = (_p(symbols), _p(indices), steps, lanes, _p(tables.cdf), tables.cdf.shape[0], tables.cdf.shape[1],
            _p(tables.length), _p(tables.offset), int(precision))
    # every push spills at most one word: symbols + escapes are bounded by (2 + 14 nibbles) pushes per symbol, but
    # real messages are far smaller -- ask the coder for the exact size first when the optimistic buffer is short
    cap = 2 * lanes + steps * lanes + 1024
    out = np.empty(cap, dtype=np.uint32)
    n = lib.hfc_rans_encode_host(*args, _p(out), cap)
    if n < 0:
        need = lib.hfc_rans_encode_host(*args, ctypes.c_void_p(0), 0)
        if need < 0:
            check(int(need), "rans_encode")
        out = np.empty(int(need), dtype=np.uint32)
        n = lib.hfc_rans_encode_host(*args, _p(out), int(need))
        if n < 0:
            check(int(n), "rans_encode")
    return out[:int(n)].copy()


def vec_ans_index_decoder(encoded, indices, tables, precision=PRECISION_P):
    """flat uint32 message + int32 indices [steps][lanes] -> int32 symbols [steps][lanes]."""
    encoded = np.ascontiguousarray(encoded, dtype=np.uint32)
    indices = _i32(indices)
    steps, lanes = indices.shape
    out = np.empty((steps, lanes), dtype=np.int32)
    check(lib.hfc_rans_decode_host(_p(encoded), encoded.shape[0], _p(indices), steps, lanes, _p(tables.cdf),
                                   tables.cdf.shape[0], tables.cdf.shape[1], _p(tables.length), _p(tables.offset),
                                   int(precision), _p(out)), "rans_decode")
    return out
