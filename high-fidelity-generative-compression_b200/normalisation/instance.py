"""InstanceNorm2D_wrap with the reference's contract (src/normalisation/instance.py:7-15): a torch.nn.InstanceNorm2d
(affine, no running statistics) whose parameters keep the reference's state_dict names (`weight`, `bias`).

Inside Encoder / Generator (use_channel_norm = False) the normalisation is executed by libhfc's `hfc_instancenorm` /
`hfc_instancenorm_bwd` on the NHWC rows of the plans (hific_b200.train_plan); this module is the parameter holder and
the stand-alone entry point for NCHW tensors.
"""
import torch
import torch.nn as nn

from .. import ops


class InstanceNorm2D(nn.InstanceNorm2d):
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("InstanceNorm2D: hific_b200 has no CPU path")
        if self.track_running_stats or not self.affine:
            raise NotImplementedError("InstanceNorm2D: the HiFIC variant is affine without running statistics")
        if abs(self.eps - ops.IN_EPS) > 1e-12:
            raise ValueError("libhfc InstanceNorm kernels are called with eps=1e-5 (torch's default, which the reference keeps)")
        n, c, h, w = x.shape
        if c % 4 != 0:
            raise NotImplementedError("stand-alone InstanceNorm2D needs channels % 4 == 0")
        rows = x.detach().permute(0, 2, 3, 1).reshape(-1, c).contiguous()
        geom = ops.Geom(n, h, w, c, ops.round_up(c, 8))
        _, out = ops.instancenorm(rows, geom, self.weight.detach(), self.bias.detach(), want_f32=True, want_act=False)
        return out.view(n, h, w, c).permute(0, 3, 1, 2).contiguous()


def InstanceNorm2D_wrap(input_channels, momentum=0.1, affine=True, track_running_stats=False, **kwargs):
    return InstanceNorm2D(input_channels, momentum=momentum, affine=affine, track_running_stats=track_running_stats)
