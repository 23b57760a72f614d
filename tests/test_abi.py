"""CPU: the C-ABI library loads and exports every symbol include/hfc.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hfc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hfc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("hfc_conv_forward", "hfc_conv_pack_weights", "hfc_conv_query", "hfc_nchw_to_act", "hfc_channelnorm",
              "hfc_latent_likelihood", "hfc_hyperlatent_likelihood", "hfc_abi_version", "hfc_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from hific_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"libhfc.so does not export {s}"
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert _lib.lib.hfc_abi_version() == 1


def test_struct_layouts_match_header():
    from hific_b200 import _lib
    assert ctypes.sizeof(_lib.ActGeom) == 9 * 4
    # hfc_conv_desc: 2 geoms + 25 int32/float fields
    assert ctypes.sizeof(_lib.ConvDesc) == 2 * 36 + 25 * 4


def test_conv_query_runs_without_gpu_and_validates():
    from hific_b200.ops import Conv, Geom, OUT_NHWC_F32, PAD_REFLECT
    from hific_b200._lib import HfcError
    g = Geom(32, 16, 16, 960, 960, 1, 1, 1, 1)
    c = Conv(g, 960, 3, pad_mode=PAD_REFLECT, pad=(1, 1, 1, 1), out_mode=OUT_NHWC_F32)
    assert c.info.block_n == 240 and c.info.n_tiles == 4 and c.info.m_tiles == 64 and c.info.k_total == 8640
    assert c.flops == pytest.approx(2.0 * 32 * 256 * 960 * 960 * 9)
    with pytest.raises(HfcError):
        Conv(Geom(1, 16, 16, 60, 60), 64, 3)  # cpad must be a multiple of 64
    with pytest.raises(HfcError):
        Conv(Geom(1, 16, 16, 64, 64), 960, 3, pad=(1, 1, 1, 1), norm=True)  # fused norm needs cout <= 256
