"""Generator with the reference's constructor and state_dict (src/network/generator.py:9-169), executed by
hific_b200.engine.GeneratorPlan.  channel_norm=False selects the InstanceNorm2d variant (generator.py:21-24, 81-84),
which runs layer by layer on the training plan (conv -> fp32 rows -> hfc_instancenorm), with and without autograd.
"""
import torch
import torch.nn as nn

from .. import engine, train_plan
from ..normalisation.channel import ChannelNorm2D
from ..normalisation.instance import InstanceNorm2D_wrap


class ResidualBlock(nn.Module):
    """One residual block (conv1, conv2, norm1, norm2 -- generator.py:9-44)."""

    def __init__(self, input_dims, kernel_size=3, stride=1, channel_norm=True, activation='relu'):
        super().__init__()
        if kernel_size != 3 or stride != 1 or activation != 'relu':
            raise NotImplementedError("hific_b200.ResidualBlock implements the HiFIC default block (3x3, stride 1, ReLU)")
        c = input_dims[1]
        norm = ChannelNorm2D if channel_norm is True else InstanceNorm2D_wrap
        kind = "channel" if channel_norm is True else "instance"
        self.conv1 = nn.Conv2d(c, c, kernel_size, stride=stride)
        self.conv2 = nn.Conv2d(c, c, kernel_size, stride=stride)
        self.norm1 = norm(c)
        self.norm2 = norm(c)
        self._plans = engine.PlanCache(lambda x: train_plan.ResidualBlockTrainPlan(x.shape[0], x.shape[2], x.shape[3],
                                                                                   x.shape[1], x.device, norm_kind=kind))

    def _apply(self, fn, *a, **k):
        self._plans.clear()
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        """generator.py:33-44 for a stand-alone call (inside Generator.forward the blocks run in the Generator's plans)."""
        engine._require_cuda(x, "ResidualBlock")
        plan = self._plans.get(x)
        if engine.wants_grad(self, x):
            return train_plan.run_training(plan, x, list(self.parameters()))
        return train_plan.run_inference(plan, x, list(self.parameters()))


class Generator(nn.Module):
    def __init__(self, input_dims, batch_size, C=16, activation='relu', n_residual_blocks=8, channel_norm=True,
                 sample_noise=False, noise_dim=32):
        super().__init__()
        if activation != 'relu':
            raise NotImplementedError("hific_b200.Generator implements the HiFIC default activation (ReLU)")
        self.channel_norm = channel_norm is True
        norm = ChannelNorm2D if self.channel_norm else InstanceNorm2D_wrap
        kind = "channel" if self.channel_norm else "instance"
        filters = list(engine.GeneratorPlan.FILTERS)
        self.C, self.n_residual_blocks = C, n_residual_blocks
        self.sample_noise, self.noise_dim = sample_noise, noise_dim
        trunk_noise = noise_dim if sample_noise else 0
        self.n_upsampling_layers = 4
        self.conv_block_init = nn.Sequential(norm(C), nn.ReflectionPad2d(1),
                                             nn.Conv2d(C, filters[0], kernel_size=(3, 3), stride=1),
                                             norm(filters[0]))
        self._trunk_noise = trunk_noise
        if sample_noise is True:
            filters[0] += noise_dim                      # generator.py:105-107: the trunk carries the noise channels too
        for m in range(n_residual_blocks):
            self.add_module(f"resblock_{m}", ResidualBlock((batch_size, filters[0], 0, 0), channel_norm=channel_norm))
        for i in range(1, 5):
            up = nn.ConvTranspose2d(filters[i - 1], filters[i], 3, stride=2, padding=1, output_padding=1)
            setattr(self, f"upconv_block{i}", nn.Sequential(up, norm(filters[i]), nn.ReLU()))
        self.conv_block_out = nn.Sequential(nn.ReflectionPad2d(3), nn.Conv2d(filters[-1], 3, kernel_size=(7, 7), stride=1))
        self._plans = engine.PlanCache(self._make_plan)
        self._train_plans = engine.PlanCache(lambda y: train_plan.GeneratorTrainPlan(
            y.shape[0], y.shape[2], y.shape[3], self.C, self.n_residual_blocks, 3, y.device, noise_dim=self._trunk_noise,
            norm_kind=kind))

    def _make_plan(self, y):
        n, _, h, w = y.shape
        return engine.GeneratorPlan(n, h, w, self.C, self.n_residual_blocks, 3, y.device, noise_dim=self._trunk_noise)

    def _apply(self, fn, *a, **k):
        self._plans.clear()
        self._train_plans.clear()
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        engine._require_cuda(x, "Generator")
        if engine.wants_grad(self, x):
            return train_plan.run_training(self._train_plans.get(x), x, list(self.parameters()))
        if not self.channel_norm:
            return train_plan.run_inference(self._train_plans.get(x), x, list(self.parameters()))
        return self._plans.get(x).run(self, x.contiguous())
