"""`Model` with the reference's constructor, attributes and return types (src/model.py:35-387) on top of the
B200 kernels.  Built so far: `compression_forward` (train / eval / EVALUATION-mode padding) and the
EVALUATION `forward` (reconstruction, q_bpp) -- i.e. the encode+decode forward path of the headline metric.
The loss / discriminator / backward half of the training step is not built yet and raises loudly.
"""
from collections import defaultdict, namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import hyperprior
from .graph import GraphedCall
from .config import ModelModes, ModelTypes
from .network import encoder, generator

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
        if getattr(args, "use_latent_mixture_model", False):
            raise NotImplementedError("HyperpriorDLMM is a non-default variant (SURVEY.md 8f), not built")
        if not hasattr(ModelTypes, self.model_type.upper()):
            raise ValueError("Invalid model_type: [{}]".format(self.model_type))
        if not hasattr(ModelModes, self.model_mode.upper()):
            raise ValueError("Invalid model_mode: [{}]".format(self.model_mode))
        self.image_dims, self.batch_size = args.image_dims, args.batch_size
        # EVALUATION mode in the reference also builds the host rANS tables; those stay with the reference.
        self.entropy_code = False
        self.Encoder = encoder.Encoder(self.image_dims, self.batch_size, C=args.latent_channels,
                                       channel_norm=args.use_channel_norm)
        self.Generator = generator.Generator(self.image_dims, self.batch_size, C=args.latent_channels,
                                             n_residual_blocks=args.n_residual_blocks,
                                             channel_norm=args.use_channel_norm, sample_noise=args.sample_noise,
                                             noise_dim=args.noise_dim)
        self.Hyperprior = hyperprior.Hyperprior(bottleneck_capacity=args.latent_channels,
                                                likelihood_type=args.likelihood_type, entropy_code=False)
        self.amortization_models = [self.Encoder, self.Generator]
        self.amortization_models.extend(self.Hyperprior.amortization_models)
        self.use_discriminator = (self.model_type == ModelTypes.COMPRESSION_GAN
                                  and self.model_mode != ModelModes.EVALUATION)
        self.Discriminator = None
        self.discriminator_steps = 0
        if self.use_discriminator:
            raise NotImplementedError("the Discriminator path is not built yet (SURVEY.md 8a rows D1-D3)")
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
        return super()._apply(fn, *a, **k)

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
        if train_generator is True:
            self.step_counter += 1
        intermediates, hyperinfo = self.compression_forward(x)
        if self.model_mode == ModelModes.EVALUATION:
            reconstruction = intermediates.reconstruction
            if self.args.normalize_input_image is True:
                reconstruction = (reconstruction + 1.) / 2.
            reconstruction = torch.clamp(reconstruction, min=0., max=1.)
            return reconstruction, intermediates.q_bpp
        raise NotImplementedError(
            "Model.forward in TRAINING/VALIDATION mode needs the loss kernels (MSE, LPIPS feature loss, GAN) "
            "and the backward kernels, which are not built yet; compression_forward() is available in all modes")
