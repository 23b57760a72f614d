"""Hyperprior entropy model with the reference's API (src/hyperprior.py:144-330): analysis / synthesis
networks run as fused sm_100a plans; the factorized and conditional likelihoods, the noise / round
quantisation, the -log2 bit sums and the straight-through latents are two fused elementwise kernels.
"""
import math
from collections import namedtuple

import torch
import torch.nn as nn

from . import engine, ops
from .compression import hyperprior_model, prior_model
from .compression.compression_utils import CompressionOutput
from .network import hyper

MIN_SCALE = 0.11
LOG_SCALES_MIN = -3.
MIN_LIKELIHOOD = 1e-9
MAX_LIKELIHOOD = 1e3
SMALL_HYPERLATENT_FILTERS = 192
LARGE_HYPERLATENT_FILTERS = 320

HyperInfo = namedtuple(
    "HyperInfo",
    "decoded "
    "latent_nbpp hyperlatent_nbpp total_nbpp latent_qbpp hyperlatent_qbpp total_qbpp",
)


class CodingModel(nn.Module):
    """Base class kept for API parity (src/hyperprior.py:43-139)."""

    def __init__(self, n_channels, min_likelihood=MIN_LIKELIHOOD, max_likelihood=MAX_LIKELIHOOD):
        super().__init__()
        self.n_channels = n_channels
        self.min_likelihood = float(min_likelihood)
        self.max_likelihood = float(max_likelihood)

    @staticmethod
    def _bits_and_bpp(log_sum, batch_size, spatial_shape):
        """hyperprior.py:80-93 with the log-likelihood sum already reduced on device (fp64)."""
        n_pixels = float(spatial_shape[0] * spatial_shape[1])
        n_bits = log_sum.to(torch.float32) / (batch_size * -math.log(2.))
        return n_bits, n_bits / n_pixels


class Hyperprior(CodingModel):
    def __init__(self, bottleneck_capacity=220, hyperlatent_filters=LARGE_HYPERLATENT_FILTERS, mode='large',
                 likelihood_type='gaussian', scale_lower_bound=MIN_SCALE, entropy_code=False,
                 vectorize_encoding=True, block_encode=True):
        super().__init__(n_channels=bottleneck_capacity)
        self.bottleneck_capacity = bottleneck_capacity
        self.scale_lower_bound = scale_lower_bound
        if mode == 'small':
            hyperlatent_filters = SMALL_HYPERLATENT_FILTERS
        if likelihood_type not in ('gaussian', 'logistic'):
            raise ValueError('Unknown likelihood model: {}'.format(likelihood_type))
        self.likelihood_type = likelihood_type
        self.analysis_net = hyper.HyperpriorAnalysis(C=bottleneck_capacity, N=hyperlatent_filters)
        self.synthesis_mu = hyper.HyperpriorSynthesis(C=bottleneck_capacity, N=hyperlatent_filters)
        self.synthesis_std = hyper.HyperpriorSynthesis(C=bottleneck_capacity, N=hyperlatent_filters)
        self.amortization_models = [self.analysis_net, self.synthesis_mu, self.synthesis_std]
        self.hyperlatent_likelihood = hyperprior_model.HyperpriorDensity(n_channels=hyperlatent_filters)
        if entropy_code is True:
            # src/hyperprior.py:183-193: integer probability tables for the host rANS coder
            self.hyperprior_entropy_model = hyperprior_model.HyperpriorEntropyModel(
                distribution=self.hyperlatent_likelihood)
            self.prior_density = prior_model.PriorDensity(n_channels=bottleneck_capacity,
                                                          scale_lower_bound=self.scale_lower_bound,
                                                          likelihood_type=likelihood_type)
            self.prior_entropy_model = prior_model.PriorEntropyModel(distribution=self.prior_density,
                                                                     min_scale=self.scale_lower_bound)
            self.index_tables = self.prior_entropy_model.scale_table_tensor
            self.vectorize_encoding = vectorize_encoding
            self.block_encode = block_encode

    def _latent_statistics(self, hyperlatents_decoded):
        """(means, raw scales) from the decoded hyper-latents; the 0.11 lower bound of src/hyperprior.py:214,250 is
        applied inside the kernels that consume the scales.  Both networks are deterministic kernels (no atomics, no
        split-K), so the encoder and the decoder derive bit-identical statistics from the same hyper-latents."""
        cur = torch.cuda.current_stream()
        side = self._side_stream(hyperlatents_decoded.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            latent_scales = self.synthesis_std(hyperlatents_decoded)
        latent_means = self.synthesis_mu(hyperlatents_decoded)
        cur.wait_stream(side)
        latent_scales.record_stream(cur)
        return latent_means.contiguous(), latent_scales.contiguous()

    def compress_forward(self, latents, spatial_shape, **kwargs):
        """src/hyperprior.py:195-247: y -> (hyper-latent message, latent message, shapes, Shannon estimates).  GPU:
        analysis / synthesis networks, symbols + table indices + bit estimates in coder order (csrc/symbols.cu);
        host: the rANS coder.  As in the reference the hyper-latents are coded, DEcoded again and the latent statistics
        derived from the decoded values, so that the decoder sees exactly the same (means, scales)."""
        engine._require_cuda(latents, "Hyperprior.compress_forward")
        with torch.no_grad():
            latents = latents.contiguous()
            hyperlatents = self.analysis_net(latents)
            hyperlatent_spatial_shape = hyperlatents.size()[2:]
            batch_shape = latents.size(0)
            hem, pem = self.hyperprior_entropy_model, self.prior_entropy_model
            hyperlatent_bits, hyperlatent_bpp, _ = hem._estimate_compression_bits(hyperlatents, spatial_shape)
            hyperlatents_encoded, hyper_coding_shape, _ = hem.compress(
                hyperlatents, vectorize=self.vectorize_encoding, block_encode=self.block_encode)
            hyperlatents_decoded, _ = hem.decompress(
                hyperlatents_encoded, batch_shape=batch_shape, broadcast_shape=hyperlatent_spatial_shape,
                coding_shape=hyper_coding_shape, vectorize=self.vectorize_encoding, block_decode=self.block_encode,
                device=latents.device)
            latent_means, latent_scales = self._latent_statistics(hyperlatents_decoded)
            latents_encoded, latent_coding_shape, _, log_sum = pem.compress(
                latents, means=latent_means, scales=latent_scales, vectorize=self.vectorize_encoding,
                block_encode=self.block_encode, return_bits=True)
            latent_bits, latent_bpp, _ = pem._bits(log_sum, batch_shape, spatial_shape)
        return CompressionOutput(
            hyperlatents_encoded=hyperlatents_encoded,
            latents_encoded=latents_encoded,
            hyperlatent_spatial_shape=hyperlatent_spatial_shape,
            spatial_shape=spatial_shape,
            hyper_coding_shape=hyper_coding_shape,
            latent_coding_shape=latent_coding_shape,
            batch_shape=batch_shape,
            hyperlatent_bits=hyperlatent_bits.item(),
            latent_bits=latent_bits.item(),
            total_bits=(hyperlatent_bits + latent_bits).item(),
            hyperlatent_bpp=hyperlatent_bpp.item(),
            latent_bpp=latent_bpp.item(),
            total_bpp=(hyperlatent_bpp + latent_bpp).item(),
        )

    def decompress_forward(self, compression_output, device):
        """src/hyperprior.py:249-274: messages -> decoded latents (N, C, H, W) on `device`."""
        co = compression_output
        with torch.no_grad():
            hyperlatents_decoded, _ = self.hyperprior_entropy_model.decompress(
                co.hyperlatents_encoded, batch_shape=co.batch_shape, broadcast_shape=co.hyperlatent_spatial_shape,
                coding_shape=co.hyper_coding_shape, vectorize=self.vectorize_encoding,
                block_decode=self.block_encode, device=device)
            latent_means, latent_scales = self._latent_statistics(hyperlatents_decoded)
            latent_spatial_shape = latent_scales.size()[2:]
            latents_decoded, _ = self.prior_entropy_model.decompress(
                co.latents_encoded, means=latent_means, scales=latent_scales, broadcast_shape=latent_spatial_shape,
                coding_shape=co.latent_coding_shape, vectorize=self.vectorize_encoding,
                block_decode=self.block_encode)
        return latents_decoded.to(device)

    def forward(self, latents, spatial_shape, **kwargs):
        engine._require_cuda(latents, "Hyperprior")
        if engine.wants_grad(self, latents):
            return self._forward_autograd(latents, spatial_shape)
        latents = latents.contiguous()
        batch = latents.shape[0]
        hyperlatents = self.analysis_net(latents)
        # sums = [ln p(noisy z), ln p(round z), ln p(noisy y), ln p(round y)], accumulated in fp64 on device
        sums = torch.zeros(4, dtype=torch.float64, device=latents.device)
        # Same RNG call as the reference (hyperprior.py:65) so seeds / patched generators line up.
        noise_z = torch.nn.init.uniform_(torch.zeros_like(hyperlatents), -0.5, 0.5)
        z_noisy, z_quant, _ = ops.hyperlatent_likelihood(
            hyperlatents, self.hyperlatent_likelihood.packed_params(), noise_z, sums=sums[0:2])
        hyperlatents_decoded = z_noisy if self.training else z_quant            # hyperprior.py:294-297
        # The two synthesis networks are independent and individually too small to fill the GPU:
        # run the scale network on a side stream (fork / join is capturable in a CUDA graph).
        cur = torch.cuda.current_stream()
        side = self._side_stream(latents.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            latent_scales = self.synthesis_std(hyperlatents_decoded)             # lower bound applied in-kernel
        latent_means = self.synthesis_mu(hyperlatents_decoded)
        cur.wait_stream(side)
        latent_scales.record_stream(cur)
        noise_y = torch.nn.init.uniform_(torch.zeros_like(latents), -0.5, 0.5)
        latents_decoded, _ = ops.latent_likelihood(latents, latent_means, latent_scales, noise_y,
                                                   self.scale_lower_bound, self.likelihood_type, sums=sums[2:4])
        return self._hyperinfo(latents_decoded, sums, batch, spatial_shape)

    def _hyperinfo(self, latents_decoded, sums, batch, spatial_shape):
        # hyperprior.py:80-93: n_bits = sum(log p) / (-ln2 * B); bpp = n_bits / n_pixels
        n_pixels = float(spatial_shape[0] * spatial_shape[1])
        bpp = sums.to(torch.float32) / (batch * -math.log(2.)) / n_pixels
        noisy_hyperlatent_bpp, quantized_hyperlatent_bpp, noisy_latent_bpp, quantized_latent_bpp = bpp.unbind(0)
        return HyperInfo(
            decoded=latents_decoded,
            latent_nbpp=noisy_latent_bpp,
            hyperlatent_nbpp=noisy_hyperlatent_bpp,
            total_nbpp=noisy_latent_bpp + noisy_hyperlatent_bpp,
            latent_qbpp=quantized_latent_bpp,
            hyperlatent_qbpp=quantized_hyperlatent_bpp,
            total_qbpp=quantized_latent_bpp + quantized_hyperlatent_bpp,
        )

    def _forward_autograd(self, latents, spatial_shape):
        """Same computation recorded for autograd: the networks through their training plans, the two likelihood
        kernels through autograd Functions with hand-written backward kernels."""
        latents = latents.contiguous()
        batch = latents.shape[0]
        hyperlatents = self.analysis_net(latents)
        noise_z = torch.nn.init.uniform_(torch.zeros_like(hyperlatents), -0.5, 0.5)
        d = self.hyperlatent_likelihood
        packed = ops.pack_density_params_autograd(*d._tensors())
        z_noisy, z_quant, sums_z = ops.HyperlatentLikelihoodFn.apply(hyperlatents.contiguous(), packed, noise_z)
        hyperlatents_decoded = z_noisy if self.training else z_quant
        latent_means = self.synthesis_mu(hyperlatents_decoded)
        latent_scales = self.synthesis_std(hyperlatents_decoded)
        noise_y = torch.nn.init.uniform_(torch.zeros_like(latents), -0.5, 0.5)
        latents_decoded, sums_y = ops.LatentLikelihoodFn.apply(latents, latent_means.contiguous(),
                                                               latent_scales.contiguous(), noise_y,
                                                               self.scale_lower_bound, self.likelihood_type)
        return self._hyperinfo(latents_decoded, torch.cat([sums_z, sums_y]), batch, spatial_shape)

    def _side_stream(self, device):
        key = (device.type, device.index)
        if getattr(self, "_side", None) is None or self._side[0] != key:
            self._side = (key, torch.cuda.Stream(device=device))
        return self._side[1]


class HyperpriorDLMM(CodingModel):
    """`-LMM` variant (src/hyperprior.py:340-458): the latents are modelled by a K-component discretised mixture whose
    logits / means / log-scales come from one synthesis network; no compress path exists for it in the reference either."""

    def __init__(self, bottleneck_capacity=64, hyperlatent_filters=LARGE_HYPERLATENT_FILTERS, mode='large',
                 likelihood_type='gaussian', scale_lower_bound=MIN_SCALE, mixture_components=4, entropy_code=False):
        super().__init__(n_channels=bottleneck_capacity)
        assert bottleneck_capacity <= 128, 'Will probably run out of memory!'        # hyperprior.py:354
        self.bottleneck_capacity = bottleneck_capacity
        self.scale_lower_bound = scale_lower_bound
        self.mixture_components = mixture_components
        if mode == 'small':
            hyperlatent_filters = SMALL_HYPERLATENT_FILTERS
        if likelihood_type not in ('gaussian', 'logistic'):
            raise ValueError('Unknown likelihood model: {}'.format(likelihood_type))
        self.likelihood_type = likelihood_type
        self.analysis_net = hyper.HyperpriorAnalysis(C=bottleneck_capacity, N=hyperlatent_filters)
        self.synthesis_DLMM_params = hyper.HyperpriorSynthesisDLMM(C=bottleneck_capacity, N=hyperlatent_filters)
        self.amortization_models = [self.analysis_net, self.synthesis_DLMM_params]
        self.hyperlatent_likelihood = hyperprior_model.HyperpriorDensity(n_channels=hyperlatent_filters)

    def forward(self, latents, spatial_shape, **kwargs):
        engine._require_cuda(latents, "HyperpriorDLMM")
        grad = engine.wants_grad(self, latents)
        latents = latents.contiguous()
        batch = latents.shape[0]
        hyperlatents = self.analysis_net(latents)
        noise_z = torch.nn.init.uniform_(torch.zeros_like(hyperlatents), -0.5, 0.5)       # hyperprior.py:65 via :409
        d = self.hyperlatent_likelihood
        if grad:
            packed = ops.pack_density_params_autograd(*d._tensors())
            z_noisy, z_quant, sums_z = ops.HyperlatentLikelihoodFn.apply(hyperlatents.contiguous(), packed, noise_z)
        else:
            z_noisy, z_quant, sums_z = ops.hyperlatent_likelihood(hyperlatents, d.packed_params(), noise_z)
        hyperlatents_decoded = z_noisy if self.training else z_quant                    # hyperprior.py:421-424
        dlmm_params = self.synthesis_DLMM_params(hyperlatents_decoded).contiguous()
        noise_y = torch.nn.init.uniform_(torch.zeros_like(latents), -0.5, 0.5)           # hyperprior.py:429
        if grad:
            decoded, sums_y = ops.DlmmLikelihoodFn.apply(latents, dlmm_params, noise_y, self.likelihood_type,
                                                         bool(self.training))
        else:
            decoded, sums_y = ops.dlmm_likelihood(latents, dlmm_params, noise_y, self.likelihood_type,
                                                  straight_through=bool(self.training))
        return Hyperprior._hyperinfo(self, decoded, torch.cat([sums_z, sums_y]), batch, spatial_shape)
