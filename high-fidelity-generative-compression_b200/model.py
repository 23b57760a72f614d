"""`Model` with the reference's constructor, attributes and return types (src/model.py:35-387) on top of the
B200 kernels: `compression_forward`, `discriminator_forward`, `compression_loss`, `GAN_loss` and `forward` in all
three model modes.  The compression model (Encoder / Hyperprior / Generator + distortion, LPIPS and rate losses) is
differentiable end to end through hand-written backward kernels (hific_b200.train_plan / grad), so the reference's
`optimize_compression_loss` (train.py:54-59) works on it; so is the Discriminator with the GAN losses
(train_plan.DiscriminatorTrainPlan: spectral-norm reparametrisation, LeakyReLU masks, the upsample + concat input), i.e.
`optimize_loss(disc_loss, disc_opt)` (train.py:49-52) and the generator's adversarial term through D into G.
"""
from collections import defaultdict, namedtuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from functools import partial

from . import hyperprior, ops
from .loss import losses
from .loss.perceptual import PerceptualLoss
from .graph import GraphedCall
from .config import ModelModes, ModelTypes
from .network import discriminator, encoder, generator

Intermediates = namedtuple("Intermediates",
                           ["input_image", "reconstruction", "latents_quantized", "n_bpp", "q_bpp"])
Disc_out = namedtuple("disc_out", ["D_real", "D_gen", "D_real_logits", "D_gen_logits"])


def pad_factor(input_image, spatial_dims, factor):
    """Reflect-pad bottom/right to a multiple of `factor` (src/helpers/utils.py:50-62).  Data movement only."""
    factor_h, factor_w = (factor, factor) if isinstance(factor, int) else factor
    h, w = spatial_dims[0], spatial_dims[1]
    pad_h = (factor_h - (h % factor_h)) % factor_h
    pad_w = (factor_w - (w % factor_w)) % factor_w
    if pad_h == 0 and pad_w == 0:
        return input_image
    return F.pad(input_image, pad=(0, pad_w, 0, pad_h), mode='reflect')


class Model(nn.Module):
    def __init__(self, args, logger, storage_train=defaultdict(list), storage_test=defaultdict(list),
                 model_mode=ModelModes.TRAINING, model_type=ModelTypes.COMPRESSION):
        super().__init__()
        self.args, self.logger = args, logger
        self.model_mode, self.model_type = model_mode, model_type
        self.log_interval = args.log_interval
        self.storage_train, self.storage_test = storage_train, storage_test
        self.step_counter = 0
        if getattr(args, "use_latent_mixture_model", False) is True:                   # src/model.py:53-54
            self.args.latent_channels = self.args.latent_channels_DLMM
        if not hasattr(ModelTypes, self.model_type.upper()):
            raise ValueError("Invalid model_type: [{}]".format(self.model_type))
        if not hasattr(ModelModes, self.model_mode.upper()):
            raise ValueError("Invalid model_mode: [{}]".format(self.model_mode))
        self.image_dims, self.batch_size = args.image_dims, args.batch_size
        # src/model.py:64-66: EVALUATION mode builds the integer probability tables of the host rANS coder
        self.entropy_code = self.model_mode == ModelModes.EVALUATION
        self.Encoder = encoder.Encoder(self.image_dims, self.batch_size, C=args.latent_channels,
                                       channel_norm=args.use_channel_norm)
        self.Generator = generator.Generator(self.image_dims, self.batch_size, C=args.latent_channels,
                                             n_residual_blocks=args.n_residual_blocks,
                                             channel_norm=args.use_channel_norm, sample_noise=args.sample_noise,
                                             noise_dim=args.noise_dim)
        if getattr(args, "use_latent_mixture_model", False) is True:                   # src/model.py:76-78 (`-LMM`)
            self.Hyperprior = hyperprior.HyperpriorDLMM(bottleneck_capacity=args.latent_channels,
                                                        likelihood_type=args.likelihood_type,
                                                        mixture_components=args.mixture_components,
                                                        entropy_code=self.entropy_code)
        else:
            self.Hyperprior = hyperprior.Hyperprior(bottleneck_capacity=args.latent_channels,
                                                    likelihood_type=args.likelihood_type,
                                                    entropy_code=self.entropy_code)
        self.amortization_models = [self.Encoder, self.Generator]
        self.amortization_models.extend(self.Hyperprior.amortization_models)
        self.use_discriminator = (self.model_type == ModelTypes.COMPRESSION_GAN
                                  and self.model_mode != ModelModes.EVALUATION)
        if self.use_discriminator is True:
            assert self.args.discriminator_steps > 0, 'Must specify nonzero training steps for D!'
            self.discriminator_steps = self.args.discriminator_steps
            self.logger.info('GAN mode enabled. Training discriminator for {} steps.'.format(self.discriminator_steps))
            self.Discriminator = discriminator.Discriminator(image_dims=self.image_dims,
                                                             context_dims=self.args.latent_dims,
                                                             C=self.args.latent_channels)
            self.gan_loss = partial(losses.gan_loss, args.gan_loss_type)
        else:
            self.discriminator_steps = 0
            self.Discriminator = None
        # As in the reference (model.py:104-105) the LPIPS network is held in a plain attribute-less container so that
        # its weights stay out of Model.parameters() / state_dict(); it follows the model across devices via _apply.
        self._lpips = [None]
        self._use_graph = False
        self._graphs = {}

    def enable_cuda_graph(self, enabled=True):
        """Replay `compression_forward` from a captured CUDA graph (one per input shape / train-eval state).
        Not part of the reference API: an opt-in for serving / benchmarking.  Weights are read through the
        packed copies made at capture time, so call it again (or disable) after changing parameters."""
        self._use_graph = bool(enabled)
        self._graphs.clear()
        return self

    def _apply(self, fn, *a, **k):
        self._graphs = {}
        if self._lpips[0] is not None:
            self._lpips[0]._apply(fn)
        return super()._apply(fn, *a, **k)

    @property
    def perceptual_loss(self):
        if self._lpips[0] is None:
            dev = next(self.parameters()).device
            self._lpips[0] = PerceptualLoss(model='net-lin', net='alex', use_gpu=dev.type == 'cuda').to(dev)
        return self._lpips[0]

    def store_loss(self, key, loss):
        assert type(loss) == float, 'Call .item() on loss before storage'
        storage = self.storage_train if self.training is True else self.storage_test
        if self.writeout is True:
            storage[key].append(loss)

    def compression_forward(self, x):
        """src/model.py:119-165."""
        if self._use_graph and x.is_cuda and not torch.is_grad_enabled():
            key = (tuple(x.shape), x.device, self.training, self.model_mode)
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = GraphedCall(self._compression_forward_impl, [x])
            inter, info = g(x)
            # fresh tensors, as the eager path returns (the graph's static outputs are overwritten on replay)
            inter = Intermediates(x, inter.reconstruction.clone(), inter.latents_quantized.clone(),
                                  inter.n_bpp.clone(), inter.q_bpp.clone())
            info = type(info)(*[t.clone() for t in info])
            return inter, info
        return self._compression_forward_impl(x)

    def _compression_forward_impl(self, x):
        image_dims = tuple(x.size()[1:])
        pad = self.model_mode == ModelModes.EVALUATION and (self.training is False)
        if pad:
            x = pad_factor(x, x.size()[2:], 2 ** self.Encoder.n_downsampling_layers)
        y = self.Encoder(x)
        if pad:
            y = pad_factor(y, y.size()[2:], 2 ** self.Hyperprior.analysis_net.n_downsampling_layers)
        hyperinfo = self.Hyperprior(y, spatial_shape=x.size()[2:])
        latents_quantized = hyperinfo.decoded
        reconstruction = self.Generator(latents_quantized)
        if self.args.normalize_input_image is True:
            reconstruction = torch.tanh(reconstruction)
        if pad:
            reconstruction = reconstruction[:, :, :image_dims[1], :image_dims[2]]
        intermediates = Intermediates(x, reconstruction, latents_quantized, hyperinfo.total_nbpp,
                                      hyperinfo.total_qbpp)
        return intermediates, hyperinfo

    def forward(self, x, train_generator=False, return_intermediates=False, writeout=True):
        """src/model.py:346-387 (EVALUATION branch)."""
        self.writeout = writeout
        losses_out = dict()
        if train_generator is True:
            self.step_counter += 1
        intermediates, hyperinfo = self.compression_forward(x)
        if self.model_mode == ModelModes.EVALUATION:
            reconstruction = intermediates.reconstruction
            if self.args.normalize_input_image is True:
                reconstruction = (reconstruction + 1.) / 2.
            reconstruction = torch.clamp(reconstruction, min=0., max=1.)
            return reconstruction, intermediates.q_bpp
        compression_model_loss = self.compression_loss(intermediates, hyperinfo)
        if self.use_discriminator is True:
            D_loss, G_loss = self.GAN_loss(intermediates, train_generator)
            weighted_G_loss = self.args.beta * G_loss
            compression_model_loss = compression_model_loss + weighted_G_loss
            losses_out['disc'] = D_loss
        losses_out['compression'] = compression_model_loss
        if (self.step_counter % self.log_interval == 1):
            self.store_loss('weighted_compression_loss', compression_model_loss.item())
        if return_intermediates is True:
            return losses_out, intermediates
        return losses_out

    # ------------------------------------------------------------------ discriminator / losses
    def discriminator_forward(self, intermediates, train_generator):
        """src/model.py:167-188 (including its pairing quirk: images are stacked [real..., gen...] while the
        context latents are repeat_interleave'd)."""
        x_gen = intermediates.reconstruction
        x_real = intermediates.input_image
        if train_generator is False:
            x_gen = x_gen.detach()
        D_in = torch.cat([x_real, x_gen], dim=0)
        latents = intermediates.latents_quantized.detach()
        latents = torch.repeat_interleave(latents, 2, dim=0)
        D_out, D_out_logits = self.Discriminator(D_in, latents)
        D_out = torch.squeeze(D_out)
        D_out_logits = torch.squeeze(D_out_logits)
        D_real, D_gen = torch.chunk(D_out, 2, dim=0)
        D_real_logits, D_gen_logits = torch.chunk(D_out_logits, 2, dim=0)
        return Disc_out(D_real, D_gen, D_real_logits, D_gen_logits)

    def distortion_loss(self, x_gen, x_real):
        """mean((255 x_gen - 255 x_real)^2) -- src/model.py:190-194, one fused reduction."""
        if torch.is_grad_enabled() and x_gen.requires_grad:
            return ops.SqDiffMeanFn.apply(x_gen, x_real, 255.)
        x_gen, x_real = x_gen.contiguous(), x_real.contiguous()
        return (ops.sqdiff_sum(x_gen, x_real, 255.) / x_gen.numel()).to(torch.float32)

    def perceptual_loss_wrapper(self, x_gen, x_real, normalize=True):
        LPIPS_loss = self.perceptual_loss.forward(x_gen, x_real, normalize=normalize)
        return torch.mean(LPIPS_loss)

    def compression_loss(self, intermediates, hyperinfo):
        """src/model.py:201-241."""
        x_real = intermediates.input_image
        x_gen = intermediates.reconstruction
        if self.args.normalize_input_image is True:
            x_real = (x_real + 1.) / 2.
            x_gen = (x_gen + 1.) / 2.
        distortion_loss = self.distortion_loss(x_gen, x_real)
        perceptual_loss = self.perceptual_loss_wrapper(x_gen, x_real, normalize=True)
        weighted_distortion = self.args.k_M * distortion_loss
        weighted_perceptual = self.args.k_P * perceptual_loss
        weighted_rate, rate_penalty = losses.weighted_rate_loss(
            self.args, total_nbpp=intermediates.n_bpp, total_qbpp=intermediates.q_bpp,
            step_counter=self.step_counter, ignore_schedule=self.args.ignore_schedule)
        weighted_R_D_loss = weighted_rate + weighted_distortion
        weighted_compression_loss = weighted_R_D_loss + weighted_perceptual
        if (self.step_counter % self.log_interval == 1):
            self.store_loss('rate_penalty', rate_penalty)
            self.store_loss('distortion', distortion_loss.item())
            self.store_loss('perceptual', perceptual_loss.item())
            self.store_loss('n_rate', intermediates.n_bpp.item())
            self.store_loss('q_rate', intermediates.q_bpp.item())
            self.store_loss('n_rate_latent', hyperinfo.latent_nbpp.item())
            self.store_loss('q_rate_latent', hyperinfo.latent_qbpp.item())
            self.store_loss('n_rate_hyperlatent', hyperinfo.hyperlatent_nbpp.item())
            self.store_loss('q_rate_hyperlatent', hyperinfo.hyperlatent_qbpp.item())
            self.store_loss('weighted_rate', weighted_rate.item())
            self.store_loss('weighted_distortion', weighted_distortion.item())
            self.store_loss('weighted_perceptual', weighted_perceptual.item())
            self.store_loss('weighted_R_D', weighted_R_D_loss.item())
            self.store_loss('weighted_compression_loss_sans_G', weighted_compression_loss.item())
        return weighted_compression_loss

    def GAN_loss(self, intermediates, train_generator=False):
        """src/model.py:244-260."""
        disc_out = self.discriminator_forward(intermediates, train_generator)
        D_loss = self.gan_loss(disc_out, mode='discriminator_loss')
        G_loss = self.gan_loss(disc_out, mode='generator_loss')
        if (self.step_counter % self.log_interval == 1):
            self.store_loss('D_gen', torch.mean(disc_out.D_gen).item())
            self.store_loss('D_real', torch.mean(disc_out.D_real).item())
            self.store_loss('disc_loss', D_loss.item())
            self.store_loss('gen_loss', G_loss.item())
            self.store_loss('weighted_gen_loss', (self.args.beta * G_loss).item())
        return D_loss, G_loss

    def compress(self, x, silent=False):
        """src/model.py:262-310: x -> Encoder -> y -> Hyperprior.compress_forward -> CompressionOutput (two rANS
        messages + shapes + Shannon estimates)."""
        assert self.model_mode == ModelModes.EVALUATION and (self.training is False), (
            f'Set model mode to {ModelModes.EVALUATION} for compression.')
        spatial_shape = tuple(x.size()[2:])
        with torch.no_grad():
            x = pad_factor(x, x.size()[2:], 2 ** self.Encoder.n_downsampling_layers)
            y = self.Encoder(x)
            y = pad_factor(y, y.size()[2:], 2 ** self.Hyperprior.analysis_net.n_downsampling_layers)
            compression_output = self.Hyperprior.compress_forward(y, spatial_shape)
        n_pixels = np.prod(spatial_shape)
        attained_hbpp = 32 * len(compression_output.hyperlatents_encoded) / n_pixels
        attained_lbpp = 32 * len(compression_output.latents_encoded) / n_pixels
        attained_bpp = 32 * ((len(compression_output.hyperlatents_encoded)
                              + len(compression_output.latents_encoded)) / n_pixels)
        if silent is False:
            self.logger.info('[ESTIMATED]')
            self.logger.info(f'BPP: {compression_output.total_bpp:.3f}')
            self.logger.info(f'HL BPP: {compression_output.hyperlatent_bpp:.3f}')
            self.logger.info(f'L BPP: {compression_output.latent_bpp:.3f}')
            self.logger.info('[ATTAINED]')
            self.logger.info(f'BPP: {attained_bpp:.3f}')
            self.logger.info(f'HL BPP: {attained_hbpp:.3f}')
            self.logger.info(f'L BPP: {attained_lbpp:.3f}')
        return compression_output

    def decompress(self, compression_output):
        """src/model.py:312-344: CompressionOutput -> reconstruction in [0, 1], cropped to the original size."""
        assert self.model_mode == ModelModes.EVALUATION and (self.training is False), (
            f'Set model mode to {ModelModes.EVALUATION} for decompression.')
        device = next(self.parameters()).device
        with torch.no_grad():
            latents_decoded = self.Hyperprior.decompress_forward(compression_output, device=device)
            reconstruction = self.Generator(latents_decoded)
            if self.args.normalize_input_image is True:
                reconstruction = torch.tanh(reconstruction)
            image_dims = compression_output.spatial_shape
            reconstruction = reconstruction[:, :, :image_dims[0], :image_dims[1]]
            if self.args.normalize_input_image is True:
                reconstruction = (reconstruction + 1.) / 2.
            reconstruction = torch.clamp(reconstruction, min=0., max=1.)
        return reconstruction
