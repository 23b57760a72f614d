"""TEST INFRASTRUCTURE: run the product's HOST glue of the compress path (shapes, coder layouts, padding, host rANS
calls, container) on a machine without a GPU by swapping every CUDA entry point it touches for the oracle's CPU
arithmetic.  Only tests import this; the product never does (it has no CPU path).  What this can and cannot show:
it pins the Python glue + the C host coder end to end against the reference's own `Model.compress` output
(tests/golden/entropy_coding.npz: model_m1 / model_m2); the CUDA kernels themselves are checked by the `-m gpu` tests.
"""
import contextlib

import numpy as np
import torch

from hific_b200 import engine, hyperprior, ops
from hific_b200._lib import SYM_PIXEL_STEPS
from hific_b200.compression import hyperprior_model
from hific_b200.network import encoder, generator, hyper
from oracle import entropy_oracle as EO
from oracle import hific_oracle as O


def _to_layout(t, layout):
    if layout == SYM_PIXEL_STEPS:
        return t.permute(0, 2, 3, 1).reshape(-1).contiguous()
    return t.reshape(-1).contiguous()


def _from_layout(flat, shape, layout):
    n, c, h, w = shape
    if layout == SYM_PIXEL_STEPS:
        return flat.reshape(n, h, w, c).permute(0, 3, 1, 2).contiguous()
    return flat.reshape(n, c, h, w)


def quantize_symbols(x, mean=None, scale_raw=None, scale_table=None, scale_lower_bound=0.11, likelihood_type="gaussian",
                     layout=0, want_symbols=True, want_indices=True, want_dequant=False, want_bits=False):
    n, c, h, w = x.shape
    out = {}
    sym = torch.floor(x + 0.5 - mean) if mean is not None else torch.floor(x + 0.5)
    sym = sym.to(torch.int32)
    if want_symbols:
        out["symbols"] = _to_layout(sym, layout)
    if want_indices:
        if scale_raw is not None:
            idx = EO.compute_indices(torch.clamp(scale_raw, scale_lower_bound), scale_table)
        else:
            idx = torch.from_numpy(EO.hyper_indices((n, c, h, w)))
        out["indices"] = _to_layout(idx, layout)
    if want_dequant:
        out["dequant"] = sym.float() + mean if mean is not None else sym.float()
    if want_bits:
        bits = EO.prior_bits(x, mean, torch.clamp(scale_raw, scale_lower_bound), likelihood_type)
        out["bits_sum"] = (bits * -np.log(2.)).to(torch.float64)
    return out


def scale_indices(scale_raw, scale_table, scale_lower_bound=0.11, layout=0):
    return _to_layout(EO.compute_indices(torch.clamp(scale_raw, scale_lower_bound), scale_table), layout)


def dequantize_symbols(symbols, mean, shape, layout=0):
    s = _from_layout(symbols, shape, layout).float()
    return s + mean if mean is not None else s


def _sd(module):
    return {k: v.detach() for k, v in module.state_dict().items()}


@contextlib.contextmanager
def cpu_emulation():
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    patch(ops, "quantize_symbols", quantize_symbols)
    patch(ops, "scale_indices", scale_indices)
    patch(ops, "dequantize_symbols", dequantize_symbols)
    patch(engine, "_require_cuda", lambda x, who: None)
    patch(encoder.Encoder, "forward", lambda self, x: O.encoder_forward(_sd(self), x, prefix=""))
    patch(generator.Generator, "forward", lambda self, y: O.generator_forward(_sd(self), y, prefix=""))
    patch(hyper.HyperpriorAnalysis, "forward", lambda self, y: O.hyper_analysis(_sd(self), y, prefix=""))
    patch(hyper.HyperpriorSynthesis, "forward", lambda self, z: O.hyper_synthesis(_sd(self), z, prefix=""))
    patch(hyperprior.Hyperprior, "_latent_statistics",
          lambda self, z: (self.synthesis_mu(z).contiguous(), self.synthesis_std(z).contiguous()))

    def hyper_bits(self, x, spatial_shape):
        params = {k: v.detach() for k, v in self.distribution.state_dict().items()}
        n_bits = EO.hyper_bits(x, params).to(torch.float32)
        return n_bits, n_bits / float(np.prod(spatial_shape)), n_bits / x.shape[0]

    patch(hyperprior_model.HyperpriorEntropyModel, "_estimate_compression_bits", hyper_bits)
    try:
        yield
    finally:
        for obj, name, value in reversed(saved):
            setattr(obj, name, value)
