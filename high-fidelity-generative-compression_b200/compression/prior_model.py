"""Conditional (mean-scale) entropy model of the latents with the reference's API
(src/compression/prior_model.py:30-305): 64 log-spaced scale rows of quantised CDFs, a table index per latent
element, symbols = floor(y + .5 - mean).  The per-element work runs in `hfc_quantize_symbols` /
`hfc_scale_indices` / `hfc_dequantize_symbols` (csrc/symbols.cu) and produces the arrays in the order the coder
walks them; the coder itself (csrc/entropy_host.cpp) runs on the host.
"""
import numpy as np
import scipy.stats
import torch
import torch.nn as nn

from .. import ops
from .._lib import SYM_BATCH_STEPS, SYM_PIXEL_STEPS
from . import compression_utils, entropy_coding, entropy_models

MIN_SCALE = entropy_models.MIN_SCALE
MIN_LIKELIHOOD = entropy_models.MIN_LIKELIHOOD
MAX_LIKELIHOOD = entropy_models.MAX_LIKELIHOOD
TAIL_MASS = entropy_models.TAIL_MASS
PRECISION_P = entropy_models.PRECISION_P

SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def prior_scale_table(scales_min=SCALES_MIN, scales_max=SCALES_MAX, levels=SCALES_LEVELS):
    """prior_model.py:25-27."""
    return torch.Tensor(np.exp(np.linspace(np.log(scales_min), np.log(scales_max), levels)))


def coder_layout(batch):
    """The vectorised coder walks batch-1 tensors pixel by pixel with the channels as lanes and larger batches
    element by element with (C, H, W) as lanes (entropy_coding.py:298-316, PATCH_SIZE (1, 1))."""
    return SYM_PIXEL_STEPS if batch == 1 else SYM_BATCH_STEPS


def coder_shape(shape):
    """(steps, lanes, coding_shape) of an (N, C, H, W) tensor."""
    n, c, h, w = shape
    if n == 1:
        return h * w, c, (c, 1, 1)
    return n, c * h * w, (c, h, w)


def coder_to_nchw(a, shape):
    n, c, h, w = shape
    a = np.asarray(a)
    return (a.T if n == 1 else a).reshape(n, c, h, w)


class PriorDensity(nn.Module):
    """Gaussian / logistic latent density convolved with U(-1/2, 1/2) (prior_model.py:250-305)."""

    def __init__(self, n_channels, min_likelihood=MIN_LIKELIHOOD, max_likelihood=MAX_LIKELIHOOD,
                 scale_lower_bound=MIN_SCALE, likelihood_type='gaussian', **kwargs):
        super().__init__()
        if likelihood_type not in ('gaussian', 'logistic'):
            raise ValueError('Unknown likelihood model: {}'.format(likelihood_type))
        self.n_channels = n_channels
        self.min_likelihood, self.max_likelihood = float(min_likelihood), float(max_likelihood)
        self.scale_lower_bound = scale_lower_bound
        self.likelihood_type = likelihood_type
        self.dtype = torch.float32
        dist = scipy.stats.norm if likelihood_type == 'gaussian' else scipy.stats.logistic
        self.standardized_quantile = dist.ppf
        self.quantile = lambda quantile, mean, scale: dist.ppf(quantile, loc=mean, scale=scale)

    def standardized_CDF(self, value):
        """maths.py:102-109 (host tensors: table building)."""
        if self.likelihood_type == 'gaussian':
            return 0.5 * torch.erfc(value * (-1. / np.sqrt(2)))
        return torch.sigmoid(value)

    def quantization_offset(self, mean, **kwargs):
        return mean.detach()

    def lower_tail(self, tail_mass, mean, scale):
        return self.quantile(0.5 * float(tail_mass), mean=mean, scale=scale)

    def upper_tail(self, tail_mass, mean, scale):
        return self.quantile(1. - 0.5 * float(tail_mass), mean=mean, scale=scale)

    def likelihood(self, x, mean, scale, **kwargs):
        """prior_model.py:292-305 with plain tensor ops (API parity / table checks; the hot path evaluates this inside
        hfc_latent_likelihood and hfc_quantize_symbols and never materialises the tensor)."""
        x = torch.abs(x - mean)
        p = self.standardized_CDF((0.5 - x) / scale) - self.standardized_CDF(-(0.5 + x) / scale)
        return torch.clamp(p, min=self.min_likelihood)

    def forward(self, x, mean, scale, **kwargs):
        return self.likelihood(x, mean, scale)


class PriorEntropyModel(entropy_models.ContinuousEntropyModel):
    def __init__(self, distribution, scale_table=None, index_ranges=64, min_scale=MIN_SCALE,
                 likelihood_bound=MIN_LIKELIHOOD, tail_mass=TAIL_MASS, precision=PRECISION_P):
        super().__init__(distribution=distribution, likelihood_bound=likelihood_bound, tail_mass=tail_mass,
                         precision=precision)
        self.index_ranges = int(index_ranges)
        self.min_scale = min_scale
        if scale_table is None:
            scale_table = prior_scale_table()
        self.scale_table = torch.clamp(torch.as_tensor(scale_table, dtype=torch.float32), min=self.min_scale)
        self.standardized_CDF = distribution.standardized_CDF
        self.standardized_quantile = distribution.standardized_quantile
        self.quantile = distribution.quantile
        self.build_tables()
        self.register_buffer('scale_table_tensor', torch.Tensor(tuple(float(s) for s in self.scale_table)))
        self.register_buffer('min_scale_tensor', torch.Tensor([float(self.min_scale)]))

    def build_tables(self, **kwargs):
        """prior_model.py:77-120.  The PMF rows are evaluated with the same host tensor ops as the reference (the
        integer tables must be identical on the encoding and the decoding machine); the quantisation of every row --
        the reference's Python double loop -- is `hfc_pmf_to_quantized_cdf_host`."""
        multiplier = -self.standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).to(torch.int32)
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = self.scale_table.unsqueeze(1).float()
        upper = self.standardized_CDF((.5 - samples) / samples_scale)
        lower = self.standardized_CDF((-.5 - samples) / samples_scale)
        pmf = (upper - lower).numpy()
        tail_mass = (2 * lower[:, :1]).numpy()
        cdf = np.zeros((len(pmf_length), max_length + 2), dtype=np.int32)
        for n in range(len(pmf_length)):
            length = int(pmf_length[n])
            row = np.concatenate((pmf[n, :length], tail_mass[n]))
            cdf[n, :length + 2] = entropy_coding.pmf_to_quantized_cdf(row, self.precision)
        self._register_tables(cdf, (-pmf_center).to(torch.int32).numpy(), (pmf_length + 2).to(torch.int32).numpy())
        compression_utils.check_argument_shapes(self.CDF, self.CDF_length, self.CDF_offset)

    # ------------------------------------------------------------------------------------------ GPU half
    def _lt(self):
        return self.distribution.likelihood_type

    def _quantize(self, x, means, scales, want_bits=False, want_symbols=True):
        n, c, h, w = x.shape
        return ops.quantize_symbols(x, means, scales, self.scale_table_tensor, SCALES_MIN, self._lt(),
                                    coder_layout(n), want_symbols=want_symbols, want_indices=want_symbols,
                                    want_bits=want_bits)

    def _estimate_compression_bits(self, x, means, scales, spatial_shape):
        """prior_model.py:122-146 -> (n_bits, bpp, bpi); one launch, the likelihood tensor is never materialised."""
        assert len(spatial_shape) == 2, 'Mispecified spatial dims'
        out = self._quantize(x, means, scales, want_bits=True, want_symbols=False)
        return self._bits(out["bits_sum"], x.shape[0], spatial_shape)

    @staticmethod
    def _bits(log_sum, batch_size, spatial_shape):
        n_bits = (log_sum / -np.log(2.)).to(torch.float32)
        return n_bits, n_bits / float(np.prod(spatial_shape)), n_bits / batch_size

    def compute_indices(self, scales):
        """prior_model.py:148-156, NCHW int32 (API parity; the coder path gets them in coder order from _quantize)."""
        n, c, h, w = scales.shape
        return ops.scale_indices(scales, self.scale_table_tensor, SCALES_MIN, SYM_BATCH_STEPS).view(n, c, h, w)

    def compress(self, bottleneck, means, scales, vectorize=True, block_encode=True, return_bits=False):
        """prior_model.py:158-198 -> (encoded uint32 message, coding_shape, rounded int32 NCHW)."""
        if not vectorize:
            raise NotImplementedError("only the reference's default coder (vectorize_encoding=True) is built")
        assert bottleneck.dim() == 4, 'Expect (N,C,H,W)-format input.'
        shape = tuple(bottleneck.shape)
        steps, lanes, coding_shape = coder_shape(shape)
        out = self._quantize(bottleneck, means, scales, want_bits=return_bits)
        symbols = out["symbols"].cpu().numpy().reshape(steps, lanes)
        indices = out["indices"].cpu().numpy().reshape(steps, lanes)
        encoded = entropy_coding.vec_ans_index_encoder(symbols, indices, self.host_tables(), self.precision)
        rounded = torch.from_numpy(np.ascontiguousarray(coder_to_nchw(symbols, shape)))
        if return_bits:
            return encoded, coding_shape, rounded, out["bits_sum"]
        return encoded, coding_shape, rounded

    def decompress(self, encoded, means, scales, broadcast_shape, coding_shape, vectorize=True, block_decode=True):
        """prior_model.py:201-246 -> (decoded fp32 NCHW on the device of `means`, raw symbols)."""
        if not vectorize:
            raise NotImplementedError("only the reference's default coder (vectorize_encoding=True) is built")
        n = scales.shape[0]
        shape = (n, self.distribution.n_channels, *broadcast_shape)
        assert tuple(scales.shape) == shape, 'Invalid indices!'
        assert means.size(1) == shape[1], 'Mean dims mismatch!'
        steps, lanes, _ = coder_shape(shape)
        layout = coder_layout(n)
        indices = ops.scale_indices(scales, self.scale_table_tensor, SCALES_MIN, layout).cpu().numpy()
        symbols = entropy_coding.vec_ans_index_decoder(encoded, indices.reshape(steps, lanes), self.host_tables(),
                                                       self.precision)
        sym_dev = torch.from_numpy(symbols).to(means.device, non_blocking=False)
        decoded = ops.dequantize_symbols(sym_dev, means.expand(shape).contiguous(), shape, layout)
        decoded_raw = torch.from_numpy(np.ascontiguousarray(coder_to_nchw(symbols, shape))).to(torch.float32)
        return decoded, decoded_raw
