"""Rate / GAN losses with the reference's function names (src/loss/losses.py:8-66)."""
import numpy as np
import torch

from .. import ops


def get_scheduled_params(param, param_schedule, step_counter, ignore_schedule=False):
    """src/helpers/utils.py:64-72."""
    if ignore_schedule is False:
        vals, steps = param_schedule['vals'], param_schedule['steps']
        assert (len(vals) == len(steps) + 1), f'Mispecified schedule! - {param_schedule}'
        idx = np.where(step_counter < np.array(steps + [step_counter + 1]))[0][0]
        param *= vals[idx]
    return param


def weighted_rate_loss(config, total_nbpp, total_qbpp, step_counter, ignore_schedule=False):
    """lambda_A if the (quantised) rate exceeds the target else lambda_B -- losses.py:8-28.  `total_qbpp.item()` is
    the one host sync per step the reference also has (losses.py:21)."""
    lambda_A = get_scheduled_params(config.lambda_A, config.lambda_schedule, step_counter, ignore_schedule)
    lambda_B = get_scheduled_params(config.lambda_B, config.lambda_schedule, step_counter, ignore_schedule)
    assert lambda_A > lambda_B, "Expected lambda_A > lambda_B, got (A) {} <= (B) {}".format(lambda_A, lambda_B)
    target_bpp = get_scheduled_params(config.target_rate, config.target_schedule, step_counter, ignore_schedule)
    rate_penalty = lambda_A if total_qbpp.item() > target_bpp else lambda_B
    return rate_penalty * total_nbpp, float(rate_penalty)


def gan_loss(gan_loss_type, disc_out, mode='generator_loss'):
    """losses.py:52-66.  The non-saturating losses (the HiFIC default) come from one fused reduction over the logits
    (`hfc_gan_sums`, backward `hfc_gan_grad`); the least-squares variant (losses.py:43-50, a non-default option on
    B*256 sigmoid outputs) is three torch reductions with torch autograd."""
    if gan_loss_type == 'least_squares':
        D_real, D_gen = disc_out.D_real, disc_out.D_gen
        if mode == 'generator_loss':
            return 0.5 * torch.mean(torch.square(D_gen - 1.0))
        return 0.5 * (torch.mean(torch.square(D_real - 1.0)) + torch.mean(torch.square(D_gen)))
    if gan_loss_type != 'non_saturating':
        raise ValueError('Invalid GAN loss')
    if torch.is_grad_enabled() and (disc_out.D_real_logits.requires_grad or disc_out.D_gen_logits.requires_grad):
        return ops.GanLossFn.apply(disc_out.D_real_logits, disc_out.D_gen_logits, 0 if mode == 'generator_loss' else 1)
    logits = torch.cat([disc_out.D_real_logits.reshape(-1), disc_out.D_gen_logits.reshape(-1)])
    n = disc_out.D_real_logits.numel()
    sums = ops.gan_sums(logits).to(torch.float32) / n          # means, as F.binary_cross_entropy_with_logits
    D_loss = sums[0] + sums[1]
    G_loss = sums[2]
    return G_loss if mode == 'generator_loss' else D_loss
