"""The LPIPS AlexNet trunk on the B200 kernels -- replaces the cuDNN execution of
torchvision's `alexnet.features` (src/loss/perceptual_similarity/pretrained_networks.py:56-94) and the per-layer
reduction of networks_basic.py:61-89, forward and backward, for both images in one batch of 2n.

  Conv2d(3, 64, 11, stride 4, pad 2)   -> 3x3 stride-1 tcgen05 conv over the 4x4 space-to-depth of the scaled image
                                          (48 channels; weights re-indexed w'[o][(dy, dx, c)][a][b] = w[o][c][4a+dy][4b+dx])
  MaxPool2d(3, 2)                      -> hfc_maxpool3s2
  Conv2d(64, 192, 5, pad 2), Conv2d(192, 384, 3, pad 1), Conv2d(384, 256, 3, pad 1), Conv2d(256, 256, 3, pad 1)
                                       -> the same conv kernel with TMA zero fill, ReLU in the epilogue
  normalize / diff / lin / spatial mean -> hfc_lpips_nhwc on the NHWC fp16 features
The trunk is frozen: the backward pass is five data-gradient convolutions (grad.ConvGrad), two max-pool adjoints and the
fused "layer gradient + incoming gradient + ReLU mask" kernel, for the reconstruction half of the batch only.
"""
import torch

from .. import ops
from ..grad import ConvGrad, GradScale
from ..ops import ACT_RELU, PAD_ZERO, Conv, Geom

_CONV_IDX = (0, 3, 6, 8, 10)           # conv layers inside torchvision's alexnet.features
_CH = (64, 192, 384, 256, 256)


def s2d_weights(w):
    """(64, 3, 11, 11) -> (64, 48, 3, 3): w'[o][(dy*4 + dx)*3 + c][a][b] = w[o][c][4a + dy][4b + dx] (zero for index 11)."""
    o = w.shape[0]
    wp = torch.zeros((o, 3, 12, 12), dtype=w.dtype, device=w.device)
    wp[:, :, :11, :11] = w
    wp = wp.view(o, 3, 3, 4, 3, 4)                       # o, c, a, dy, b, dx
    return wp.permute(0, 3, 5, 1, 2, 4).reshape(o, 48, 3, 3).contiguous()


class LpipsTrunkPlan:
    def __init__(self, n, h, w, device):
        self.n, self.h, self.w = n, h, w
        n2 = 2 * n
        oh1, ow1 = (h + 4 - 11) // 4 + 1, (w + 4 - 11) // 4 + 1
        if oh1 < 7 or ow1 < 7:
            raise ValueError("LPIPS trunk: image too small for AlexNet's two max-pools")
        self.hs, self.ws = oh1 + 2, ow1 + 2
        pool = lambda v: (v - 3) // 2 + 1
        p1h, p1w = pool(oh1), pool(ow1)
        p2h, p2w = pool(p1h), pool(p1w)
        G = lambda nn, hh, ww, c, cp=None: Geom(nn, hh, ww, c, cp or c)
        self.g_s2d = G(n2, self.hs, self.ws, 48, 64)
        self.g_f = [G(n2, oh1, ow1, 64), G(n2, p1h, p1w, 192), G(n2, p2h, p2w, 384), G(n2, p2h, p2w, 256),
                    G(n2, p2h, p2w, 256)]
        self.g_p = [G(n2, p1h, p1w, 64), G(n2, p2h, p2w, 192)]
        ins = [self.g_s2d, self.g_p[0], self.g_p[1], self.g_f[2], self.g_f[3]]
        self.ks, self.pads = (3, 5, 3, 3, 3), (0, 2, 1, 1, 1)
        self.convs = [Conv(ins[i], _CH[i], self.ks[i], pad_mode=PAD_ZERO, pad=(self.pads[i],) * 4, out_geom=self.g_f[i],
                           act=ACT_RELU) for i in range(5)]
        self.s2d = self.g_s2d.alloc(device)
        self.feats = [g.alloc(device) for g in self.g_f]
        self.pools = [g.alloc(device) for g in self.g_p]
        self.flops = sum(c.flops for c in self.convs)
        # backward: reconstruction half only
        half = lambda g: Geom(n, g.h, g.w, g.c, g.cpad)
        self.h_f = [half(g) for g in self.g_f]
        self.grads = [ConvGrad(half(ins[i]), _CH[i], self.ks[i], stride=1, pad_mode=PAD_ZERO, pad=(self.pads[i],) * 4)
                      for i in range(5)]
        self._w1, self._w1_key = None, None

    def weights(self, trunk):
        ws = [trunk[i].weight for i in _CONV_IDX]
        key = (ws[0].data_ptr(), ws[0]._version, ws[0].device)
        if self._w1 is None or self._w1_key != key:
            self._w1, self._w1_key = s2d_weights(ws[0].detach()), key
        return [self._w1] + [w.detach() for w in ws[1:]], [trunk[i].bias.detach() for i in _CONV_IDX]

    def forward(self, owner, target, pred, normalize):
        """-> (n,) fp32 LPIPS distances; keeps the feature maps for backward()."""
        ws, bs = self.weights(owner.trunk)
        ops.lpips_prep(target, pred, self.g_s2d, normalize, owner.shift, owner.scale, out=self.s2d)
        x = self.s2d
        for i in range(5):
            x = self.convs[i](x, ws[i], bs[i], out=self.feats[i])
            if i < 2:
                x = ops.maxpool3s2(x, self.g_f[i], self.g_p[i], out=self.pools[i])
        out = torch.zeros(self.n, dtype=torch.float32, device=pred.device)
        for i in range(5):
            ops.lpips_nhwc(self.feats[i], self.g_f[i], owner.lins[i], out)
        return out

    def backward(self, owner, upstream, normalize):
        """upstream: (n,) d L / d distance -> d L / d pred (n, 3, h, w)."""
        ws, _ = self.weights(owner.trunk)
        n = self.n
        g_in = None
        for i in (4, 3, 2, 1, 0):
            g = ops.lpips_nhwc_bwd(self.feats[i], self.g_f[i], owner.lins[i], upstream, g_in)
            dx = self.grads[i].data_grad(g, ws[i])                   # rows over the conv's input pixels
            if i in (2, 1):                                          # the conv's input is a max-pooled feature map
                g_in = ops.maxpool3s2_bwd(dx, self.feats[i - 1][n:], self.h_f[i - 1])
            else:
                g_in = dx
        return ops.lpips_prep_bwd(g_in, n, self.h, self.w, self.hs, self.ws, normalize, owner.scale)


class LpipsTrunkFn(torch.autograd.Function):
    """(pred, target) -> per-image LPIPS distance through the native trunk; gradient w.r.t. pred only."""

    @staticmethod
    def forward(ctx, pred, target, plan, owner, normalize):
        ctx.plan, ctx.owner, ctx.normalize = plan, owner, normalize
        return plan.forward(owner, target, pred, normalize)

    @staticmethod
    def backward(ctx, d_out):
        plan = ctx.plan
        if not hasattr(plan, "grad_scale"):
            plan.grad_scale = GradScale()
        dpred, _ = plan.grad_scale.run(
            lambda d: (plan.backward(ctx.owner, d.to(torch.float32).contiguous(), ctx.normalize), []), d_out)
        return dpred, None, None, None, None
