"""CPU oracle for the HiFIC forward hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, stateless restatement (torch CPU, fp32) of what the reference
(Justin-Tan/high-fidelity-generative-compression @ 7d4e9e7) computes on the path
Encoder -> Hyperprior (analysis / factorized + conditional likelihood / synthesis) -> Generator
(-> Discriminator, losses).  Every function cites the reference file:line it follows.  Parameters
arrive as a flat dict using the reference's own ``state_dict`` key names.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module, and only as the checker / the reported CPU baseline.  The product
package (``hific_b200``) must never import it.

PARITY PINNING: this restatement is pinned against the real reference modules imported from
``/root/reference`` (see ``oracle/make_golden.py``); the resulting golden vectors are committed in
``tests/golden/`` and ``tests/test_oracle_golden.py`` replays them on every CPU test run.
"""
import math
from collections import namedtuple

import torch
import torch.nn.functional as F

MIN_SCALE = 0.11          # src/hyperprior.py:12
MIN_LIKELIHOOD = 1e-9     # src/hyperprior.py:14
CN_EPS = 1e-3             # src/normalisation/channel.py:35

HyperOut = namedtuple(
    "HyperOut",
    "hyperlatents noisy_hyperlatents quantized_hyperlatents latent_means latent_scales decoded "
    "latent_nbpp hyperlatent_nbpp total_nbpp latent_qbpp hyperlatent_qbpp total_qbpp")


def _ident(t):
    return t


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def channel_norm(x, gamma, beta, eps=CN_EPS):
    """ChannelNorm2D.forward -- src/normalisation/channel.py:48-59 (unbiased variance over dim 1)."""
    mu = torch.mean(x, dim=1, keepdim=True)
    var = torch.var(x, dim=1, keepdim=True)
    return gamma * ((x - mu) * torch.rsqrt(var + eps)) + beta


def _conv(sd, prefix, x, stride=1, rnd=_ident):
    return F.conv2d(rnd(x), rnd(sd[prefix + ".weight"]), sd[prefix + ".bias"], stride=stride)


def _convT(sd, prefix, x, stride, padding, output_padding, rnd=_ident):
    return F.conv_transpose2d(rnd(x), rnd(sd[prefix + ".weight"]), sd[prefix + ".bias"], stride=stride,
                              padding=padding, output_padding=output_padding)


IN_EPS = 1e-5


def _cn(sd, prefix, x):
    """The inter-layer normalisation: ChannelNorm2D (gamma / beta in the state dict), or -- use_channel_norm = False --
    torch.nn.InstanceNorm2d(affine=True, track_running_stats=False) (weight / bias; src/normalisation/instance.py:7-15,
    encoder.py:41-44, generator.py:21-24, 81-84)."""
    if prefix + ".gamma" in sd:
        return channel_norm(x, sd[prefix + ".gamma"], sd[prefix + ".beta"])
    return F.instance_norm(x, weight=sd[prefix + ".weight"], bias=sd[prefix + ".bias"], eps=IN_EPS)


def _reflect(x, pad):
    """nn.ReflectionPad2d(pad) with pad = (left, right, top, bottom)."""
    return F.pad(x, pad, mode="reflect")


# ------------------------------------------------------------------------------------------------
# Encoder -- src/network/encoder.py:56-111
# ------------------------------------------------------------------------------------------------
def encoder_forward(sd, x, prefix="Encoder.", rnd=_ident, taps=None):
    p = prefix
    # conv_block1: ReflectionPad2d(3), Conv 7x7 s1, ChannelNorm, ReLU            (encoder.py:56-61)
    h = F.relu(_cn(sd, p + "conv_block1.2", _conv(sd, p + "conv_block1.1", _reflect(x, (3, 3, 3, 3)), rnd=rnd)))
    if taps is not None:
        taps["E1"] = h
    # conv_block2..5: ReflectionPad2d((0,1,1,0)), Conv 3x3 s2, ChannelNorm, ReLU  (encoder.py:64-93)
    for i in range(2, 6):
        h = _reflect(h, (0, 1, 1, 0))
        h = F.relu(_cn(sd, p + f"conv_block{i}.2", _conv(sd, p + f"conv_block{i}.1", h, stride=2, rnd=rnd)))
        if taps is not None:
            taps[f"E{i}"] = h
    # conv_block_out: ReflectionPad2d(1), Conv 3x3 s1                            (encoder.py:98-101)
    return _conv(sd, p + "conv_block_out.1", _reflect(h, (1, 1, 1, 1)), rnd=rnd)


# ------------------------------------------------------------------------------------------------
# Generator -- src/network/generator.py:33-44, 98-169
# ------------------------------------------------------------------------------------------------
def residual_block(sd, prefix, x, rnd=_ident):
    """ResidualBlock.forward -- generator.py:33-44."""
    res = _conv(sd, prefix + ".conv1", _reflect(x, (1, 1, 1, 1)), rnd=rnd)
    res = F.relu(_cn(sd, prefix + ".norm1", res))
    res = _conv(sd, prefix + ".conv2", _reflect(res, (1, 1, 1, 1)), rnd=rnd)
    res = _cn(sd, prefix + ".norm2", res)
    return res + x


def generator_forward(sd, y_hat, n_residual_blocks=9, prefix="Generator.", rnd=_ident, taps=None, noise=None):
    """`noise`: the (B, noise_dim, H, W) draw of the `sample_noise=True` variant (generator.py:149-153), concatenated to
    the head; None = the default architecture."""
    p = prefix
    # conv_block_init: ChannelNorm, ReflectionPad2d(1), Conv 3x3, ChannelNorm    (generator.py:98-103)
    head = _cn(sd, p + "conv_block_init.0", y_hat)
    head = _conv(sd, p + "conv_block_init.2", _reflect(head, (1, 1, 1, 1)), rnd=rnd)
    head = _cn(sd, p + "conv_block_init.3", head)
    if taps is not None:
        taps["G0"] = head
    if noise is not None:
        head = torch.cat((head, noise.to(head)), dim=1)
    x = head
    for m in range(n_residual_blocks):                                           # generator.py:154-159
        x = residual_block(sd, p + f"resblock_{m}", x, rnd=rnd)
        if taps is not None:
            taps[f"R{m}"] = x
    x = x + head                                                                 # generator.py:161
    for i in range(1, 5):                                                        # generator.py:115-137
        x = _convT(sd, p + f"upconv_block{i}.0", x, 2, 1, 1, rnd=rnd)
        x = F.relu(_cn(sd, p + f"upconv_block{i}.1", x))
        if taps is not None:
            taps[f"U{i}"] = x
    # conv_block_out: ReflectionPad2d(3), Conv 7x7                               (generator.py:139-142)
    return _conv(sd, p + "conv_block_out.1", _reflect(x, (3, 3, 3, 3)), rnd=rnd)


# ------------------------------------------------------------------------------------------------
# Hyperprior networks -- src/network/hyper.py:52-63, 83-97
# ------------------------------------------------------------------------------------------------
def hyper_analysis(sd, y, prefix="Hyperprior.analysis_net.", rnd=_ident):
    p = prefix
    h = F.relu(F.conv2d(rnd(y), rnd(sd[p + "conv1.weight"]), sd[p + "conv1.bias"], stride=1, padding=1))
    h = F.relu(_conv(sd, p + "conv2", _reflect(h, (2, 2, 2, 2)), stride=2, rnd=rnd))   # padding_mode='reflect'
    return _conv(sd, p + "conv3", _reflect(h, (2, 2, 2, 2)), stride=2, rnd=rnd)


def hyper_synthesis(sd, z, prefix, rnd=_ident):
    h = F.relu(_convT(sd, prefix + "conv1", z, 2, 2, 1, rnd=rnd))
    h = F.relu(_convT(sd, prefix + "conv2", h, 2, 2, 1, rnd=rnd))
    return _convT(sd, prefix + "conv3", h, 1, 1, 0, rnd=rnd)


# ------------------------------------------------------------------------------------------------
# Factorized density -- src/compression/hyperprior_model.py:305-326, 349-384
# ------------------------------------------------------------------------------------------------
def density_cdf_logits(sd, x, prefix="Hyperprior.hyperlatent_likelihood."):
    """x: (C, 1, *).  Note the tanh gate is applied after all four layers (hyperprior_model.py:324)."""
    logits = x
    for k in range(4):
        H, a, b = sd[prefix + f"H_{k}"], sd[prefix + f"a_{k}"], sd[prefix + f"b_{k}"]
        logits = torch.bmm(F.softplus(H), logits) + b
        logits = logits + torch.tanh(a) * torch.tanh(logits)
    return logits


class _LowerBoundToward(torch.autograd.Function):
    """LowerBoundToward -- src/helpers/maths.py:87-100: forward clamp(min=bound); backward passes the gradient where
    the input was >= bound OR the gradient is negative (i.e. would move the input up, towards the bound)."""

    @staticmethod
    def forward(ctx, tensor, bound):
        ctx.mask = tensor.ge(bound)
        return torch.clamp(tensor, bound)

    @staticmethod
    def backward(ctx, grad_output):
        gate = torch.logical_or(ctx.mask, grad_output.lt(0.)).type(grad_output.dtype)
        return grad_output * gate, None


def lower_bound(x, bound):
    return _LowerBoundToward.apply(x, bound)


def density_likelihood(sd, z, prefix="Hyperprior.hyperlatent_likelihood."):
    """HyperpriorDensity.likelihood -- hyperprior_model.py:349-384 (N,C,H,W in and out)."""
    n, c, h, w = z.shape
    lat = z.permute(1, 0, 2, 3).reshape(c, 1, -1)
    upper = density_cdf_logits(sd, lat + 0.5, prefix)
    lower = density_cdf_logits(sd, lat - 0.5, prefix)
    sign = -torch.sign(upper + lower)
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    lik = lower_bound(lik, MIN_LIKELIHOOD)
    return lik.reshape(c, n, h, w).permute(1, 0, 2, 3)


# ------------------------------------------------------------------------------------------------
# CodingModel -- src/hyperprior.py:57-139, src/helpers/maths.py:102-109
# ------------------------------------------------------------------------------------------------
def standardized_cdf(v, likelihood_type="gaussian"):
    if likelihood_type == "gaussian":
        return 0.5 * torch.erfc(v * (-1.0 / math.sqrt(2.0)))     # maths.py:102-105
    return torch.sigmoid(v)                                       # maths.py:107-109


def latent_likelihood(x, mean, scale, likelihood_type="gaussian"):
    """CodingModel.latent_likelihood -- hyperprior.py:124-139."""
    d = torch.abs(x - mean)
    upper = standardized_cdf((0.5 - d) / scale, likelihood_type)
    lower = standardized_cdf(-(0.5 + d) / scale, likelihood_type)
    return lower_bound(upper - lower, MIN_LIKELIHOOD)


def estimate_entropy(likelihood, spatial_shape):
    """CodingModel._estimate_entropy -- hyperprior.py:80-93. Returns (n_bits, bpp)."""
    batch = likelihood.shape[0]
    n_pixels = spatial_shape[0] * spatial_shape[1]
    n_bits = torch.sum(torch.log(likelihood + 1e-9)) / (batch * -math.log(2.0))
    return n_bits, n_bits / n_pixels


def quantize_st(x, mean):
    """CodingModel.quantize_latents_st -- hyperprior.py:108-122 (forward value)."""
    v = x - mean
    delta = (torch.floor(v + 0.5) - v).detach()      # "ignore rounding in backward pass" (hyperprior.py:116)
    return (v + delta) + mean


def hyperprior_forward(sd, y, spatial_shape, training, noise_z=None, noise_y=None, likelihood_type="gaussian",
                       prefix="Hyperprior.", rnd=_ident):
    """Hyperprior.forward -- src/hyperprior.py:277-330.

    noise_z / noise_y replace the two ``torch.nn.init.uniform_(-0.5, 0.5)`` draws (hyperprior.py:65), consumed
    in that order; ``None`` means zeros (only meaningful for the eval branch, which ignores the noisy terms)."""
    p = prefix
    z = hyper_analysis(sd, y, p + "analysis_net.", rnd=rnd)
    nz = z + (noise_z if noise_z is not None else torch.zeros_like(z))                 # mode='noise'
    _, hl_nbpp = estimate_entropy(density_likelihood(sd, nz, p + "hyperlatent_likelihood."), spatial_shape)
    qz = torch.floor(z + 0.5)                                                        # mode='quantize', no means
    _, hl_qbpp = estimate_entropy(density_likelihood(sd, qz, p + "hyperlatent_likelihood."), spatial_shape)
    z_dec = nz if training else qz                                                   # hyperprior.py:294-297
    mu = hyper_synthesis(sd, z_dec, p + "synthesis_mu.", rnd=rnd)
    sigma = lower_bound(hyper_synthesis(sd, z_dec, p + "synthesis_std.", rnd=rnd), MIN_SCALE)
    ny = y + (noise_y if noise_y is not None else torch.zeros_like(y))
    _, l_nbpp = estimate_entropy(latent_likelihood(ny, mu, sigma, likelihood_type), spatial_shape)
    qy = torch.floor(y - mu + 0.5) + mu                                              # hyperprior.py:68-71
    _, l_qbpp = estimate_entropy(latent_likelihood(qy, mu, sigma, likelihood_type), spatial_shape)
    decoded = quantize_st(y, mu)
    return HyperOut(z, nz, qz, mu, sigma, decoded, l_nbpp, hl_nbpp, l_nbpp + hl_nbpp, l_qbpp, hl_qbpp,
                    l_qbpp + hl_qbpp)


# ------------------------------------------------------------------------------------------------
# Model.compression_forward -- src/model.py:119-165 ; utils.pad_factor -- src/helpers/utils.py:50-62
# ------------------------------------------------------------------------------------------------
def pad_factor(x, factor):
    h, w = x.shape[2:]
    ph = (factor - h % factor) % factor
    pw = (factor - w % factor) % factor
    return F.pad(x, (0, pw, 0, ph), mode="reflect")


def compression_forward(sd, x, training, evaluation_mode=False, noise_z=None, noise_y=None, n_residual_blocks=9,
                        likelihood_type="gaussian", rnd=_ident):
    """Returns (reconstruction, HyperOut, y).  ``evaluation_mode`` mirrors ModelModes.EVALUATION with
    ``self.training is False``: pad to multiples of 16 / 4, crop afterwards (model.py:133-144,160)."""
    image_hw = x.shape[2:]
    pad = evaluation_mode and not training
    if pad:
        x = pad_factor(x, 16)
    y = encoder_forward(sd, x, rnd=rnd)
    if pad:
        y = pad_factor(y, 4)
    hyper = hyperprior_forward(sd, y, x.shape[2:], training, noise_z, noise_y, likelihood_type, rnd=rnd)
    recon = generator_forward(sd, hyper.decoded, n_residual_blocks, rnd=rnd)
    if pad:
        recon = recon[:, :, :image_hw[0], :image_hw[1]]
    return recon, hyper, y


# ------------------------------------------------------------------------------------------------
# Losses -- src/model.py:190-194, src/loss/losses.py:8-41, src/helpers/utils.py:64-72
# ------------------------------------------------------------------------------------------------
def distortion_loss(x_gen, x_real):
    return torch.mean((x_gen * 255.0 - x_real * 255.0) ** 2)


def scheduled(param, schedule, step, ignore_schedule=False):
    if ignore_schedule:
        return param
    vals, steps = schedule["vals"], schedule["steps"]
    idx = 0
    while idx < len(steps) and step >= steps[idx]:
        idx += 1
    return param * vals[idx]


def weighted_rate_loss(cfg, total_nbpp, total_qbpp, step, ignore_schedule=False):
    lam_a = scheduled(cfg["lambda_A"], cfg["lambda_schedule"], step, ignore_schedule)
    lam_b = scheduled(cfg["lambda_B"], cfg["lambda_schedule"], step, ignore_schedule)
    target = scheduled(cfg["target_rate"], cfg["target_schedule"], step, ignore_schedule)
    penalty = lam_a if float(total_qbpp.detach()) > target else lam_b
    return penalty * total_nbpp, float(penalty)


def gan_losses_non_saturating(d_real_logits, d_gen_logits):
    """_non_saturating_loss -- losses.py:30-41. Returns (D_loss, G_loss)."""
    bce = F.binary_cross_entropy_with_logits
    d_loss = bce(d_real_logits, torch.ones_like(d_real_logits)) + bce(d_gen_logits, torch.zeros_like(d_gen_logits))
    g_loss = bce(d_gen_logits, torch.ones_like(d_gen_logits))
    return d_loss, g_loss


# ------------------------------------------------------------------------------------------------
# Discriminator -- src/network/discriminator.py:35-86 (spectral norm: torch.nn.utils.spectral_norm)
# ------------------------------------------------------------------------------------------------
def spectral_normalize(w_orig, u, v=None, n_power_iterations=1, eps=1e-12, training=True):
    """torch.nn.utils.spectral_norm hook semantics (discriminator.py:46-62): in training mode run one power
    iteration from the stored u (updating u, v), then W = W_orig / sigma.  Returns (W, u_new, v_new)."""
    w_mat = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():     # torch's hook iterates under no_grad: u, v are constants of the backward pass
            for _ in range(n_power_iterations):
                v = F.normalize(torch.mv(w_mat.t(), u), dim=0, eps=eps)
                u = F.normalize(torch.mv(w_mat, v), dim=0, eps=eps)
    sigma = torch.dot(u, torch.mv(w_mat, v))
    return w_orig / sigma, u, v


def discriminator_forward(sd, x, y, prefix="Discriminator.", training=True, rnd=_ident):
    """Discriminator.forward -- discriminator.py:66-86.  Returns (sigmoid(logits), logits, new_uv dict)."""
    p = prefix
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    c = lrelu(_conv(sd, p + "context_conv", _reflect(y, (1, 1, 1, 1)), rnd=rnd))
    c = F.interpolate(c, scale_factor=16, mode="nearest")
    h = torch.cat((x, c), dim=1)
    new_uv = {}
    for i in range(1, 5):
        w, u, v = spectral_normalize(sd[p + f"conv{i}.weight_orig"], sd[p + f"conv{i}.weight_u"],
                                     sd[p + f"conv{i}.weight_v"], training=training)
        new_uv[f"conv{i}"] = (u, v)
        h = lrelu(F.conv2d(rnd(_reflect(h, (1, 1, 1, 1))), rnd(w), sd[p + f"conv{i}.bias"], stride=2))
    logits = _conv(sd, p + "conv_out", h, rnd=rnd).reshape(-1, 1)
    return torch.sigmoid(logits), logits, new_uv


# ------------------------------------------------------------------------------------------------
# LPIPS -- src/loss/perceptual_similarity/networks_basic.py:61-98, perceptual_loss.py:26-46
# ------------------------------------------------------------------------------------------------
LPIPS_SLICES = ((0, 2), (2, 5), (5, 8), (8, 10), (10, 12))   # pretrained_networks.py:67-76


def lpips_forward(trunk_features, lin_weights, pred, target, normalize=True):
    """trunk_features: torchvision alexnet().features (frozen); lin_weights: 5 tensors (C_k,).
    Returns (N,1,1,1) like PerceptualLoss.forward(pred, target, normalize)."""
    if normalize:
        target, pred = 2 * target - 1, 2 * pred - 1
    shift = torch.tensor([-.030, -.088, -.188], device=pred.device).view(1, 3, 1, 1)
    scale = torch.tensor([.458, .448, .450], device=pred.device).view(1, 3, 1, 1)

    def feats(x):
        h = (x - shift) / scale
        outs = []
        for lo, hi in LPIPS_SLICES:
            for i in range(lo, hi):
                h = trunk_features[i](h)
            outs.append(h)
        return outs

    f0, f1 = feats(target), feats(pred)          # DistModel.forward(in0=target, in1=pred)
    val = 0
    for k in range(5):
        n0 = f0[k] / torch.sqrt(torch.sum(f0[k] ** 2, dim=1, keepdim=True) + 1e-10)
        n1 = f1[k] / torch.sqrt(torch.sum(f1[k] ** 2, dim=1, keepdim=True) + 1e-10)
        d = (n0 - n1) ** 2
        val = val + (d * lin_weights[k].view(1, -1, 1, 1)).sum(dim=1, keepdim=True).mean(dim=(2, 3), keepdim=True)
    return val


# ------------------------------------------------------------------------------------------------
# One COMPRESSION_GAN training forward -- Model.forward src/model.py:346-387 with compression_loss :201-241,
# GAN_loss :244-260, discriminator_forward :167-188.  Returns (compression_loss, disc_loss, new_uv); call
# .backward() on the one train.py would (train.py:137-141: compression on generator steps, disc otherwise).
# ------------------------------------------------------------------------------------------------
def gan_training_losses(sd, x, noise_z, noise_y, cfg, lpips_trunk, lpips_lins, train_generator, step=1,
                        n_residual_blocks=9):
    recon, hyper, _ = compression_forward(sd, x, True, False, noise_z, noise_y, n_residual_blocks=n_residual_blocks)
    dist = distortion_loss(recon, x)
    lp = lpips_forward(lpips_trunk, lpips_lins, recon, x).mean()
    rate, _ = weighted_rate_loss(cfg, hyper.total_nbpp, hyper.total_qbpp, step)
    x_gen = recon if train_generator else recon.detach()
    d_in = torch.cat([x, x_gen], 0)
    lat = torch.repeat_interleave(hyper.decoded.detach(), 2, dim=0)
    _, logits, new_uv = discriminator_forward(sd, d_in, lat, training=True)
    d_real, d_gen = torch.chunk(logits.squeeze(), 2, dim=0)
    d_loss, g_loss = gan_losses_non_saturating(d_real, d_gen)
    comp = rate + cfg["k_M"] * dist + cfg["k_P"] * lp + cfg["beta"] * g_loss
    return comp, d_loss, new_uv


# ------------------------------------------------------------------------------------------------
# helpers for precision studies (not part of the reference): emulate 16-bit operand rounding
# ------------------------------------------------------------------------------------------------
def round_fp16(t):
    return t.half().float()


def round_bf16(t):
    return t.bfloat16().float()


# ------------------------------------------------------------------------------------------------
# HyperpriorDLMM -- src/hyperprior.py:340-458 with HyperpriorSynthesisDLMM src/network/hyper.py:100-130 and
# unpack_likelihood_params hyper.py:19-35 (discretised mixture likelihood of the latents; `-LMM` in train.py:227).
# Pinned against the real reference by tests/golden/dlmm_c8.npz (oracle/make_golden_dlmm.py).
# ------------------------------------------------------------------------------------------------
LOG_SCALES_MIN = -3.0


def hyper_synthesis_dlmm(sd, z, prefix="Hyperprior.synthesis_DLMM_params.", rnd=_ident):
    """HyperpriorSynthesisDLMM.forward -- hyper.py:121-130: two ReLU transposed convs, a linear transposed conv and a
    linear 1x1 conv to 3*K*C channels."""
    h = F.relu(_convT(sd, prefix + "conv1", z, 2, 2, 1, rnd))
    h = F.relu(_convT(sd, prefix + "conv2", h, 2, 2, 1, rnd))
    h = _convT(sd, prefix + "conv3", h, 1, 1, 0, rnd)
    return F.conv2d(rnd(h), rnd(sd[prefix + "conv_out.weight"]), sd[prefix + "conv_out.bias"])


def dlmm_log_likelihood(x, dlmm_params, likelihood_type="gaussian"):
    """HyperpriorDLMM.latent_log_likelihood_DLMM -- hyperprior.py:379-401.  x (N, C, H, W), dlmm_params (N, 3*C*K, H, W)
    -> log-likelihood (N, C, H, W)."""
    n, c, h, w = x.shape
    k = dlmm_params.shape[1] // (3 * c)
    p = dlmm_params.reshape(n, 3, c, k, h, w)
    logit_pis, means = p[:, 0], p[:, 1]
    log_scales = lower_bound(p[:, 2], LOG_SCALES_MIN)
    xc = torch.abs(x.reshape(n, c, 1, h, w) - means)
    inv_stds = torch.exp(-log_scales)
    upper = standardized_cdf(inv_stds * (0.5 - xc), likelihood_type)
    lower = standardized_cdf(inv_stds * (-0.5 - xc), likelihood_type)
    pmf = lower_bound(upper - lower, MIN_LIKELIHOOD)
    return torch.logsumexp(F.log_softmax(logit_pis, dim=2) + torch.log(pmf), dim=2)


def estimate_entropy_log(log_likelihood, spatial_shape):
    """CodingModel._estimate_entropy_log -- hyperprior.py:95-106."""
    batch = log_likelihood.shape[0]
    n_bits = torch.sum(log_likelihood) / (batch * -math.log(2.0))
    return n_bits, n_bits / (spatial_shape[0] * spatial_shape[1])


def hyperprior_dlmm_forward(sd, y, spatial_shape, training, noise_z=None, noise_y=None, likelihood_type="gaussian",
                            prefix="Hyperprior.", rnd=_ident):
    """HyperpriorDLMM.forward -- hyperprior.py:403-458 (noise tensors as in hyperprior_forward)."""
    p = prefix
    z = hyper_analysis(sd, y, p + "analysis_net.", rnd=rnd)
    nz = z + (noise_z if noise_z is not None else torch.zeros_like(z))
    _, hl_nbpp = estimate_entropy(density_likelihood(sd, nz, p + "hyperlatent_likelihood."), spatial_shape)
    qz = torch.floor(z + 0.5)
    _, hl_qbpp = estimate_entropy(density_likelihood(sd, qz, p + "hyperlatent_likelihood."), spatial_shape)
    z_dec = nz if training else qz
    params = hyper_synthesis_dlmm(sd, z_dec, p + "synthesis_DLMM_params.", rnd=rnd)
    ny = y + (noise_y if noise_y is not None else torch.zeros_like(y))
    _, l_nbpp = estimate_entropy_log(dlmm_log_likelihood(ny, params, likelihood_type), spatial_shape)
    qy = torch.floor(y + 0.5)
    _, l_qbpp = estimate_entropy_log(dlmm_log_likelihood(qy, params, likelihood_type), spatial_shape)
    decoded = (y + (torch.floor(y + 0.5) - y).detach()) if training else qy          # hyperprior.py:443-446
    return HyperOut(z, nz, qz, params, None, decoded, l_nbpp, hl_nbpp, l_nbpp + hl_nbpp, l_qbpp, hl_qbpp,
                    l_qbpp + hl_qbpp)
