"""Encoder with the reference's constructor and state_dict (src/network/encoder.py:9-111), executed by the
fused sm_100a plan in hific_b200.engine (tcgen05 implicit-GEMM convs, ChannelNorm/ReLU/reflection-pad fused).
channel_norm=False selects the InstanceNorm2d variant (encoder.py:41-44): it runs layer by layer (conv -> fp32 rows ->
hfc_instancenorm) on the training plan, with and without autograd.
"""
import torch
import torch.nn as nn

from .. import engine, train_plan
from ..normalisation.channel import ChannelNorm2D
from ..normalisation.instance import InstanceNorm2D_wrap


class Encoder(nn.Module):
    def __init__(self, image_dims, batch_size, activation='relu', C=220, channel_norm=True):
        super().__init__()
        if activation != 'relu':
            raise NotImplementedError("hific_b200.Encoder implements the HiFIC default activation (ReLU)")
        self.channel_norm = channel_norm is True
        norm = ChannelNorm2D if self.channel_norm else InstanceNorm2D_wrap
        kind = "channel" if self.channel_norm else "instance"
        filters = engine.EncoderPlan.FILTERS
        self.im_channels, self.C = image_dims[0], C
        self.n_downsampling_layers = 4
        # Containers only hold parameters under the reference's names: conv_block{i}.1.{weight,bias},
        # conv_block{i}.2.{gamma,beta}; index 0 is the (parameter-free) ReflectionPad2d, index 3 the ReLU.
        cin = self.im_channels
        for i, cout in enumerate(filters):
            pad = nn.ReflectionPad2d(3) if i == 0 else nn.ReflectionPad2d((0, 1, 1, 0))
            conv = nn.Conv2d(cin, cout, kernel_size=7 if i == 0 else 3, stride=1 if i == 0 else 2)
            setattr(self, f"conv_block{i + 1}", nn.Sequential(pad, conv, norm(cout), nn.ReLU()))
            cin = cout
        self.conv_block_out = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(cin, C, 3, stride=1))
        self._plans = engine.PlanCache(self._make_plan)
        self._train_plans = engine.PlanCache(lambda x: train_plan.EncoderTrainPlan(
            x.shape[0], x.shape[2], x.shape[3], self.im_channels, self.C, x.device, norm_kind=kind))

    def _make_plan(self, x):
        n, _, h, w = x.shape
        return engine.EncoderPlan(n, h, w, self.im_channels, self.C, x.device)

    def _apply(self, fn, *a, **k):
        self._plans.clear()          # buffers and packed weights live on the old device
        self._train_plans.clear()
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        engine._require_cuda(x, "Encoder")
        if engine.wants_grad(self, x):
            return train_plan.run_training(self._train_plans.get(x), x, list(self.parameters()))
        if not self.channel_norm:
            return train_plan.run_inference(self._train_plans.get(x), x, list(self.parameters()))
        return self._plans.get(x).run(self, x.contiguous())
