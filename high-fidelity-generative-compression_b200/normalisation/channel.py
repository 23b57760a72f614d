"""ChannelNorm2D with the reference's parameter contract (src/normalisation/channel.py:29-59).

Inside Encoder / Generator the norm is fused into the producing convolution's epilogue (or runs as the
row kernel `hfc_channelnorm` for 480/960 channels); this module is the parameter holder and the stand-alone
entry point for NCHW tensors.
"""
import torch
import torch.nn as nn

from .. import ops


class ChannelNorm2D(nn.Module):
    def __init__(self, input_channels, momentum=0.1, eps=1e-3, affine=True, **kwargs):
        super().__init__()
        if abs(eps - ops.CN_EPS) > 1e-12:
            raise ValueError("libhfc ChannelNorm kernels are built for eps=1e-3 (the reference default)")
        if not affine:
            raise NotImplementedError("ChannelNorm2D(affine=False) is not on the HiFIC path")
        self.momentum, self.eps, self.affine = momentum, eps, affine
        self.gamma = nn.Parameter(torch.ones(1, input_channels, 1, 1))
        self.beta = nn.Parameter(torch.zeros(1, input_channels, 1, 1))

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("ChannelNorm2D: hific_b200 has no CPU path")
        n, c, h, w = x.shape
        if c % 4 != 0:
            raise NotImplementedError("stand-alone ChannelNorm2D needs channels % 4 == 0")
        rows = x.permute(0, 2, 3, 1).reshape(-1, c).contiguous()
        geom = ops.Geom(n, h, w, c, ops.round_up(c, 8))
        _, out = ops.channelnorm(rows, geom, self.gamma, self.beta, want_f32=True, want_act=False)
        return out.view(n, h, w, c).permute(0, 3, 1, 2).contiguous()


def ChannelNorm2D_wrap(input_channels, momentum=0.1, affine=True, track_running_stats=False, **kwargs):
    return ChannelNorm2D(input_channels, momentum=momentum, affine=affine)
