"""Patch discriminator with the reference's constructor and state_dict (src/network/discriminator.py:8-86):
`context_conv`, spectrally normalised `conv1..conv4` (`weight_orig`, `weight_u`, `weight_v`), `conv_out`.
Executed by hific_b200.engine.DiscriminatorPlan: tcgen05 convs with fused bias + LeakyReLU(0.2), the
upsample + concat folded into one layout kernel, spectral-norm power iteration as small GEMV kernels whose
1/sigma is folded into the weight packing.
"""
import torch
import torch.nn as nn

from .. import engine


class Discriminator(nn.Module):
    def __init__(self, image_dims, context_dims, C, spectral_norm=True):
        super().__init__()
        if not spectral_norm:
            raise NotImplementedError("weight_norm variant is not on the HiFIC path")
        self.image_dims, self.context_dims = image_dims, context_dims
        im_channels = image_dims[0]
        self.im_channels, self.C = im_channels, C
        filters = engine.DiscriminatorPlan.FILTERS
        cnn_kwargs = dict(stride=2, padding=1, padding_mode='reflect')
        # parameter containers with the reference's names / init order; nn.utils.spectral_norm registers
        # weight_orig (parameter) and weight_u / weight_v (buffers) exactly as in the reference
        self.context_conv = nn.Conv2d(C, engine.DiscriminatorPlan.CONTEXT_C, kernel_size=3, padding=1, padding_mode='reflect')
        self.context_upsample = nn.Upsample(scale_factor=16, mode='nearest')
        self.activation = nn.LeakyReLU(negative_slope=0.2)
        cin = im_channels + engine.DiscriminatorPlan.CONTEXT_C
        for i, f in enumerate(filters):
            setattr(self, f"conv{i + 1}", nn.utils.spectral_norm(nn.Conv2d(cin, f, 4, **cnn_kwargs)))
            cin = f
        self.conv_out = nn.Conv2d(filters[3], 1, kernel_size=1, stride=1)
        self._plans = engine.PlanCache(self._make_plan)

    def _make_plan(self, key_tensor):
        n, _, h, w, ch, cw = self._plan_key
        return engine.DiscriminatorPlan(n, h, w, self.im_channels, self.C, ch, cw, key_tensor.device)

    def _apply(self, fn, *a, **k):
        self._plans.clear()
        return super()._apply(fn, *a, **k)

    def forward(self, x, y):
        engine._require_cuda(x, "Discriminator")
        engine._require_cuda(y, "Discriminator")
        if x.shape[0] != y.shape[0]:
            raise ValueError("Discriminator: image and context batch sizes differ")
        # the plan cache is keyed on a tensor's shape: fold both shapes into a dummy key
        self._plan_key = (x.shape[0], x.shape[1], x.shape[2], x.shape[3], y.shape[2], y.shape[3])
        key = torch.empty((0,) + tuple(self._plan_key), device=x.device)
        plan = self._plans.get(key)
        if engine.wants_grad(self, x, y):
            if y.requires_grad:
                raise NotImplementedError("Discriminator: no gradient path into the context latents (the reference "
                                          "detaches them, src/model.py:178)")
            from .. import train_plan
            if getattr(plan, "train", None) is None:
                plan.train = train_plan.DiscriminatorTrainPlan(plan)
            params = [self.context_conv.weight, self.context_conv.bias]
            for i in range(1, 5):
                layer = getattr(self, f"conv{i}")
                params += [layer.weight_orig, layer.bias]
            params += [self.conv_out.weight, self.conv_out.bias]
            logits = train_plan.DiscriminatorFunction.apply(plan.train, self, x, y, *params)
        else:
            logits = plan.run(self, x.contiguous(), y.contiguous())
        out_logits = logits.view(-1, 1)
        return torch.sigmoid(out_logits), out_logits
