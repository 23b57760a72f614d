"""Factorized hyper-latent density (parameter contract of the reference's HyperpriorDensity,
src/compression/hyperprior_model.py:252-387).  The likelihood itself is evaluated by the fused kernel
`hfc_hyperlatent_likelihood` (see hific_b200.hyperprior.Hyperprior.forward); this module owns the
parameters H_k, a_k, b_k and their packed (softplus / tanh pre-applied) device copy.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .._lib import SYM_BATCH_STEPS
from . import compression_utils, entropy_coding, entropy_models
from .prior_model import coder_layout, coder_shape, coder_to_nchw

MIN_SCALE = entropy_models.MIN_SCALE
MIN_LIKELIHOOD = 1e-9
MAX_LIKELIHOOD = 1e3
TAIL_MASS = entropy_models.TAIL_MASS
PRECISION_P = entropy_models.PRECISION_P


class HyperpriorEntropyModel(entropy_models.ContinuousEntropyModel):
    """Table-driven coder of the hyper-latents (src/compression/hyperprior_model.py:21-248): one quantised CDF per
    channel, symbols = floor(z + .5)."""

    def __init__(self, distribution, likelihood_bound=MIN_LIKELIHOOD, tail_mass=TAIL_MASS, precision=PRECISION_P):
        super().__init__(distribution=distribution, likelihood_bound=likelihood_bound, tail_mass=tail_mass,
                         precision=precision)

    def compute_medians(self):
        self.medians = self.distribution.median().view(1, -1, 1, 1).cpu()

    def build_tables(self, **kwargs):
        """hyperprior_model.py:42-105.  Host-side and init-time: tails by `estimate_tails`, PMF rows by the density
        in plain tensor ops on the CPU (encoder and decoder must derive identical integers from the checkpoint), row
        quantisation by `hfc_pmf_to_quantized_cdf_host`."""
        offsets = 0.
        density = self.distribution.host_copy()
        lower_tail = density.lower_tail(self.tail_mass)
        upper_tail = density.upper_tail(self.tail_mass)
        self.medians = density.median().view(1, -1, 1, 1)
        minima = torch.clamp(torch.ceil(offsets - lower_tail).to(torch.int32), min=0)
        maxima = torch.clamp(torch.ceil(upper_tail - offsets).to(torch.int32), min=0)
        pmf_start = offsets - minima.to(torch.float32)
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max())
        samples = torch.arange(max_length, dtype=torch.float32).view(1, -1) + pmf_start.view(-1, 1, 1)
        with torch.no_grad():
            pmf = density.likelihood(samples, collapsed_format=True).squeeze(1)
        cdf = np.zeros((len(pmf_length), max_length + 2), dtype=np.int32)
        for n in range(len(pmf_length)):
            length = int(pmf_length[n])
            row = pmf[n, :length]
            overflow = torch.clamp(1. - torch.sum(row, dim=0, keepdim=True), min=0.)
            cdf[n, :length + 2] = entropy_coding.pmf_to_quantized_cdf(torch.cat((row, overflow), dim=0).numpy(),
                                                                      self.precision)
        self._register_tables(cdf, (-minima).to(torch.int32).numpy(), (pmf_length + 2).to(torch.int32).numpy())
        if self.CDF.device != self.distribution.H_0.device:
            for name in ("CDF", "CDF_offset", "CDF_length"):
                getattr(self, name).data = getattr(self, name).data.to(self.distribution.H_0.device)
        compression_utils.check_argument_shapes(self.CDF, self.CDF_length, self.CDF_offset)

    def _estimate_compression_bits(self, x, spatial_shape):
        """hyperprior_model.py:108-133 -> (n_bits, bpp, bpi): the rounded branch of the fused density kernel."""
        assert len(spatial_shape) == 2, 'Mispecified spatial dims'
        _, _, sums = ops.hyperlatent_likelihood(x.contiguous(), self.distribution.packed_params(), None)
        n_bits = (sums[1] / -np.log(2.)).to(torch.float32)
        return n_bits, n_bits / float(np.prod(spatial_shape)), n_bits / x.shape[0]

    def compute_indices(self, broadcast_shape):
        """hyperprior_model.py:135-139."""
        index_size = self.distribution.n_channels
        indices = torch.arange(index_size, dtype=torch.int32).view(-1, 1, 1)
        return indices.repeat(1, *broadcast_shape)

    def _coder_indices(self, shape):
        n, c, h, w = shape
        steps, lanes, _ = coder_shape(shape)
        ch = np.arange(c, dtype=np.int32)
        if n == 1:
            return np.broadcast_to(ch[None, :], (steps, lanes))
        return np.broadcast_to(np.repeat(ch, h * w)[None, :], (steps, lanes))

    def compress(self, bottleneck, block_encode=True, vectorize=True):
        """hyperprior_model.py:141-197 -> (encoded, coding_shape, rounded int32 NCHW)."""
        if not vectorize:
            raise NotImplementedError("only the reference's default coder (vectorize_encoding=True) is built")
        assert bottleneck.dim() == 4, 'Expect (N,C,H,W)-format input.'
        shape = tuple(bottleneck.shape)
        steps, lanes, coding_shape = coder_shape(shape)
        out = ops.quantize_symbols(bottleneck.contiguous(), layout=coder_layout(shape[0]), want_indices=False)
        symbols = out["symbols"].cpu().numpy().reshape(steps, lanes)
        encoded = entropy_coding.vec_ans_index_encoder(symbols, self._coder_indices(shape), self.host_tables(),
                                                       self.precision)
        rounded = torch.from_numpy(np.ascontiguousarray(coder_to_nchw(symbols, shape)))
        return encoded, coding_shape, rounded

    def decompress(self, encoded, batch_shape, broadcast_shape, coding_shape, vectorize=True, block_decode=True,
                   device=None):
        """hyperprior_model.py:200-248 -> (decoded fp32 NCHW, raw symbols)."""
        if not vectorize:
            raise NotImplementedError("only the reference's default coder (vectorize_encoding=True) is built")
        shape = (batch_shape, self.distribution.n_channels, *broadcast_shape)
        symbols = entropy_coding.vec_ans_index_decoder(encoded, self._coder_indices(shape), self.host_tables(),
                                                       self.precision)
        device = device if device is not None else self.distribution.H_0.device
        sym_dev = torch.from_numpy(symbols).to(device)
        decoded = ops.dequantize_symbols(sym_dev, None, shape, coder_layout(batch_shape))
        decoded_raw = torch.from_numpy(np.ascontiguousarray(coder_to_nchw(symbols, shape))).to(torch.float32)
        return decoded, decoded_raw


class HyperpriorDensity(nn.Module):
    def __init__(self, n_channels, init_scale=10., filters=(3, 3, 3), min_likelihood=MIN_LIKELIHOOD,
                 max_likelihood=MAX_LIKELIHOOD, **kwargs):
        super().__init__()
        if tuple(filters) != (3, 3, 3):
            raise NotImplementedError("the fused density kernel is built for filters=(3, 3, 3)")
        self.init_scale = float(init_scale)
        self.filters = tuple(int(f) for f in filters)
        self.min_likelihood, self.max_likelihood = float(min_likelihood), float(max_likelihood)
        self.n_channels = n_channels
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for k in range(len(self.filters) + 1):
            H = nn.Parameter(torch.full((n_channels, f[k + 1], f[k]), float(np.log(np.expm1(1 / scale / f[k + 1])))))
            a = nn.Parameter(torch.zeros((n_channels, f[k + 1], 1)))
            b = nn.Parameter(torch.zeros((n_channels, f[k + 1], 1)))
            torch.nn.init.uniform_(b, -0.5, 0.5)
            self.register_parameter(f"H_{k}", H)
            self.register_parameter(f"a_{k}", a)
            self.register_parameter(f"b_{k}", b)
        self._packed, self._packed_key = None, None

    def _tensors(self):
        return ([getattr(self, f"H_{k}") for k in range(4)], [getattr(self, f"a_{k}") for k in range(4)],
                [getattr(self, f"b_{k}") for k in range(4)])

    def packed_params(self):
        Hs, a_s, bs = self._tensors()
        key = tuple((t.data_ptr(), t._version) for t in Hs + a_s + bs)
        if self._packed is None or key != self._packed_key:
            self._packed, self._packed_key = ops.pack_density_params(Hs, a_s, bs), key
        return self._packed

    # -- plain tensor-op evaluation: init-time table building on the host (and API parity); never on the hot path -----
    def host_copy(self):
        """Detached fp32 CPU copy of this density (tails / PMF rows of the coder tables are computed on the host)."""
        cp = HyperpriorDensity(self.n_channels, init_scale=self.init_scale, filters=self.filters,
                               min_likelihood=self.min_likelihood, max_likelihood=self.max_likelihood)
        cp.load_state_dict({k: v.detach().float().cpu() for k, v in self.state_dict().items()})
        return cp

    def cdf_logits(self, x, update_parameters=True):
        """hyperprior_model.py:305-326; x: (C, 1, *)."""
        logits = x
        for k in range(len(self.filters) + 1):
            H_k, a_k, b_k = getattr(self, f"H_{k}"), getattr(self, f"a_{k}"), getattr(self, f"b_{k}")
            if update_parameters is False:
                H_k, a_k, b_k = H_k.detach(), a_k.detach(), b_k.detach()
            logits = torch.bmm(F.softplus(H_k), logits)
            logits = logits + b_k
            logits = logits + torch.tanh(a_k) * torch.tanh(logits)
        return logits

    def quantization_offset(self, **kwargs):
        return 0.

    def _tail(self, target):
        f = lambda x: self.cdf_logits(x, update_parameters=False)
        t = compression_utils.estimate_tails(f, target=target, shape=torch.Size((self.n_channels, 1, 1))).detach()
        return t.reshape(self.n_channels)

    def lower_tail(self, tail_mass):
        return self._tail(-np.log(2. / tail_mass - 1.))

    def upper_tail(self, tail_mass):
        return self._tail(np.log(2. / tail_mass - 1.))

    def median(self):
        return self._tail(0.)

    def likelihood(self, x, collapsed_format=False, **kwargs):
        """hyperprior_model.py:349-384 in tensor ops (tables / API parity).  Training and evaluation use the fused
        kernel `hfc_hyperlatent_likelihood` instead."""
        latents = x
        if collapsed_format is False:
            latents = latents.permute(1, 0, 2, 3)
            shape = latents.shape
            latents = torch.reshape(latents, (shape[0], 1, -1))
        cdf_upper = self.cdf_logits(latents + 0.5)
        cdf_lower = self.cdf_logits(latents - 0.5)
        sign = -torch.sign(cdf_upper + cdf_lower).detach()
        likelihood_ = torch.abs(torch.sigmoid(sign * cdf_upper) - torch.sigmoid(sign * cdf_lower))
        likelihood_ = torch.clamp(likelihood_, min=self.min_likelihood)
        if collapsed_format is True:
            return likelihood_
        return torch.reshape(likelihood_, shape).permute(1, 0, 2, 3)

    def forward(self, x, **kwargs):
        raise NotImplementedError(
            "HyperpriorDensity is evaluated inside Hyperprior.forward by the fused likelihood kernel; the "
            "per-element likelihood tensor is never materialised on the B200 path (use .likelihood() for a "
            "tensor-op evaluation)")
