"""Geometry-bound execution plans for the HiFIC networks.

A plan is built once per (batch, height, width, device) and owns every intermediate buffer, so a
forward pass is a fixed sequence of libhfc launches with no allocation (CUDA-graph friendly).
Between layers activations stay in the internal NHWC fp16 format; NCHW fp32 appears only at the
module boundaries the reference exposes (x, y, z, mu, sigma, y_hat, x_hat).

Layer lists follow the reference networks:
  Encoder            src/network/encoder.py:56-111
  Generator          src/network/generator.py:33-44, 98-169
  HyperpriorAnalysis src/network/hyper.py:52-63
  HyperpriorSynthesis src/network/hyper.py:83-97
"""
import os

import torch

from . import ops
from .ops import (ACT_LEAKY02, ACT_NONE, ACT_RELU, OUT_NCHW_F32, OUT_NHWC_F16, OUT_NHWC_F32, PAD_REFLECT, PAD_ZERO, Conv,
                  Geom, round_up)


def _require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f"{who}: hific_b200 has no CPU path; move the module and its input to a B200 (cuda) device")
    if x.dtype != torch.float32:
        raise TypeError(f"{who}: expected float32 input, got {x.dtype}")


class PlanCache:
    """Per-module cache of plans keyed by input geometry."""

    def __init__(self, factory, max_entries=4):
        self.factory, self.max_entries, self.plans = factory, max_entries, {}

    def get(self, x):
        key = (tuple(x.shape), x.device)
        plan = self.plans.get(key)
        if plan is None:
            if len(self.plans) >= self.max_entries:
                self.plans.pop(next(iter(self.plans)))
            plan = self.plans[key] = self.factory(x)
        return plan

    def clear(self):
        self.plans.clear()


class EncoderPlan:
    FILTERS = (60, 120, 240, 480, 960)

    def __init__(self, n, h, w, im_channels, C, device):
        f = self.FILTERS
        self.g_in = Geom(n, h, w, im_channels, 8, 3, 3, 3, 4)          # ReflectionPad2d(3) + 1 spare column
        self.x_act = self.g_in.alloc(device)
        asym = (1, 0, 0, 1)                                            # ReflectionPad2d((0,1,1,0)): top 1, right 1
        convs = []
        g1 = Geom(n, h, w, f[0], 64, *asym)
        convs.append(Conv(self.g_in, f[0], 7, pad_mode=PAD_REFLECT, pad=(3, 3, 3, 3), window=True,
                          out_geom=g1, out_reflect=True, act=ACT_RELU, norm=True))
        g = g1
        self.cn = {}
        for i in range(1, 5):
            oh, ow = (g.h + 1 - 3) // 2 + 1, (g.w + 1 - 3) // 2 + 1
            border = asym if i < 4 else (1, 1, 1, 1)                   # conv_block_out uses ReflectionPad2d(1)
            g_next = Geom(n, oh, ow, f[i], round_up(f[i], 64), *border)
            if f[i] <= 256:
                convs.append(Conv(g, f[i], 3, stride=2, pad_mode=PAD_REFLECT, pad=(1, 0, 0, 1), out_geom=g_next,
                                  out_reflect=True, act=ACT_RELU, norm=True))
            else:  # 480 / 960 channels: raw fp32 rows + stand-alone ChannelNorm
                rows = Geom(n, oh, ow, f[i], f[i])
                convs.append(Conv(g, f[i], 3, stride=2, pad_mode=PAD_REFLECT, pad=(1, 0, 0, 1),
                                  out_mode=OUT_NHWC_F32, out_geom=rows))
                self.cn[i] = g_next
            g = g_next
        convs.append(Conv(g, C, 3, pad_mode=PAD_REFLECT, pad=(1, 1, 1, 1), out_mode=OUT_NCHW_F32))
        self.convs = convs
        self.bufs = [c.alloc_out(device) for c in convs[:-1]]
        self.cn_bufs = {i: g.alloc(device) for i, g in self.cn.items()}
        self.flops = sum(c.flops for c in convs)

    def run(self, mod, x):
        ops.nchw_to_act(x, self.g_in, reflect=True, out=self.x_act)
        h = self.x_act
        for i in range(5):
            blk = getattr(mod, f"conv_block{i + 1}")
            conv, norm = blk[1], blk[2]
            if i in self.cn:
                rows = self.convs[i](h, conv.weight, conv.bias, out=self.bufs[i])
                h, _ = ops.channelnorm(rows, self.cn[i], norm.gamma, norm.beta, act=ACT_RELU, reflect=True,
                                       out_act=self.cn_bufs[i])
            else:
                h = self.convs[i](h, conv.weight, conv.bias, norm.gamma, norm.beta, out=self.bufs[i])
        last = mod.conv_block_out[1]
        return self.convs[5](h, last.weight, last.bias)              # fresh NCHW fp32 tensor (module output)


class GeneratorPlan:
    FILTERS = (960, 480, 240, 120, 60)

    def __init__(self, n, h, w, C, n_residual_blocks, im_channels, device, noise_dim=0):
        f = self.FILTERS
        self.n_res = n_residual_blocks
        # sample_noise (generator.py:105-107, 149-153): noise_dim channels of N(0, 1) noise are concatenated to the
        # 960-channel head, and the residual trunk + the first transposed conv are F0 = 960 + noise_dim channels wide
        self.noise_dim = noise_dim
        F0 = f[0] + noise_dim
        self.F0 = F0
        F0p = round_up(F0, 64)
        b1 = (1, 1, 1, 1)
        self.g_in = Geom(n, h, w, C, round_up(C, 64), *b1)
        self.in_act = self.g_in.alloc(device)
        rows960 = Geom(n, h, w, 960, 960)
        rows_t = Geom(n, h, w, F0, F0)
        self.g_head = Geom(n, h, w, 960, 960, *b1)
        self.g_b1 = Geom(n, h, w, F0, F0p, *b1)       # bordered (input of a 3x3 reflect conv)
        self.g_flat = Geom(n, h, w, F0, F0p)          # border-less (input of the first transposed conv)
        self.conv_init = Conv(self.g_in, 960, 3, pad_mode=PAD_REFLECT, pad=b1, out_mode=OUT_NHWC_F32, out_geom=rows960)
        # one Conv object per residual conv so each keeps its own packed weights
        self.res_convs = [(Conv(self.g_b1, F0, 3, pad_mode=PAD_REFLECT, pad=b1, out_mode=OUT_NHWC_F32, out_geom=rows_t),
                           Conv(self.g_b1, F0, 3, pad_mode=PAD_REFLECT, pad=b1, out_mode=OUT_NHWC_F32, out_geom=rows_t))
                          for _ in range(n_residual_blocks)]
        # default (HFC_FUSE_RESNORM=0 restores the two-launch plan): every residual conv runs fused with the ChannelNorm
        # (+ReLU / + residual adds) behind it (hfc_conv_forward_widenorm: the 960-channel row of a pixel is normalised
        # across a 4-CTA cluster), which removes the stand-alone ChannelNorm launch and the fp32 round trip of the conv
        # output.  Geometries the fused kernel does not take (ragged maps, the 992-channel noise variant) use two launches.
        self.fused = None
        if os.environ.get("HFC_FUSE_RESNORM", "1") == "1" and n_residual_blocks > 0 and F0 == F0p:
            fused = []
            for i in range(n_residual_blocks):
                last = i == n_residual_blocks - 1
                c1 = Conv(self.g_b1, F0, 3, pad_mode=PAD_REFLECT, pad=b1, out_geom=self.g_b1, out_reflect=True, act=ACT_RELU)
                c2 = Conv(self.g_b1, F0, 3, pad_mode=PAD_REFLECT, pad=b1, out_geom=self.g_flat if last else self.g_b1,
                          out_reflect=not last, act=ACT_NONE)
                fused.append((c1, c2))
            if all(c.widenorm_supported() for pair in fused for c in pair):
                self.fused = fused
        m = n * h * w
        self.rows = torch.empty((m, F0), dtype=torch.float32, device=device)
        self.head_f32 = torch.empty((m, F0), dtype=torch.float32, device=device)
        self.x_f32 = [torch.empty((m, F0), dtype=torch.float32, device=device) for _ in range(2)]
        self.act_a = self.g_b1.alloc(device)
        self.act_b = self.g_b1.alloc(device)
        self.act_flat = self.g_flat.alloc(device)
        # upsampling path
        self.ups = []
        g = self.g_flat
        for i in range(1, 5):
            oh, ow = g.h * 2, g.w * 2
            border = (0, 0, 0, 0) if i < 4 else (3, 3, 3, 3)           # conv_block_out: ReflectionPad2d(3)
            g_next = Geom(n, oh, ow, f[i], round_up(f[i], 64), *border)
            if f[i] <= 256:
                conv = Conv(g, f[i], 3, stride=2, transposed=True, pad=(1, 1, 1, 1), out_geom=g_next,
                            out_reflect=any(border), act=ACT_RELU, norm=True)
                self.ups.append((conv, None, conv.alloc_out(device), None))
            else:
                rows = Geom(n, oh, ow, f[i], f[i])
                conv = Conv(g, f[i], 3, stride=2, transposed=True, pad=(1, 1, 1, 1), out_mode=OUT_NHWC_F32, out_geom=rows)
                self.ups.append((conv, g_next, conv.alloc_out(device), g_next.alloc(device)))
            g = g_next
        self.conv_out = Conv(g, im_channels, 7, pad_mode=PAD_REFLECT, pad=(3, 3, 3, 3), out_mode=OUT_NCHW_F32)
        self.flops = (self.conv_init.flops + sum(a.flops + b.flops for a, b in self.res_convs) +
                      sum(u[0].flops for u in self.ups) + self.conv_out.flops)

    def run(self, mod, y_hat):
        init = mod.conv_block_init
        ops.nchw_to_act(y_hat, self.g_in, reflect=True, norm=True, gamma=init[0].gamma, beta=init[0].beta,
                        out=self.in_act)
        if self.noise_dim == 0:
            self.conv_init(self.in_act, init[2].weight, init[2].bias, out=self.rows)
            ops.channelnorm(self.rows, self.g_b1, init[3].gamma, init[3].beta, act=ACT_NONE, reflect=True,
                            want_f32=True, out_f32=self.head_f32, out_act=self.act_a)
        else:
            # head = cat(norm(conv(.)), z), z ~ N(0, 1) drawn with torch.randn as the reference does (generator.py:150-153):
            # pure data movement on a (B, 992, 16, 16) tensor, done with torch ops
            n, hh, ww = self.g_b1.n, self.g_b1.h, self.g_b1.w
            rows960 = self.conv_init(self.in_act, init[2].weight, init[2].bias)
            _, head960 = ops.channelnorm(rows960, self.g_head, init[3].gamma, init[3].beta, act=ACT_NONE, reflect=False,
                                         want_f32=True, want_act=False)
            z = torch.randn((n, self.noise_dim, hh, ww)).to(head960)
            head = torch.cat((head960.view(n, hh, ww, 960).permute(0, 3, 1, 2), z), dim=1).contiguous()
            self.head_f32.copy_(head.permute(0, 2, 3, 1).reshape(-1, self.F0))
            ops.nchw_to_act(head, self.g_b1, reflect=True, out=self.act_a)
        x_f32, x_act = self.head_f32, self.act_a
        for m in range(self.n_res if self.fused is not None else 0):
            blk = getattr(mod, f"resblock_{m}")
            c1, c2 = self.fused[m]
            last = m == self.n_res - 1
            nxt = self.x_f32[m % 2]
            c1.call_widenorm(x_act, blk.conv1.weight, blk.conv1.bias, blk.norm1.gamma, blk.norm1.beta, out_act=self.act_b)
            c2.call_widenorm(self.act_b, blk.conv2.weight, blk.conv2.bias, blk.norm2.gamma, blk.norm2.beta, res1=x_f32,
                             res2=self.head_f32 if last else None, out_f32=None if last else nxt,
                             out_act=self.act_flat if last else self.act_a)
            x_f32, x_act = nxt, (self.act_flat if last else self.act_a)
        for m in range(self.n_res if self.fused is None else 0):
            blk = getattr(mod, f"resblock_{m}")
            c1, c2 = self.res_convs[m]
            c1(x_act, blk.conv1.weight, blk.conv1.bias, out=self.rows)
            ops.channelnorm(self.rows, self.g_b1, blk.norm1.gamma, blk.norm1.beta, act=ACT_RELU, reflect=True,
                            out_act=self.act_b)
            c2(self.act_b, blk.conv2.weight, blk.conv2.bias, out=self.rows)
            last = m == self.n_res - 1
            nxt = self.x_f32[m % 2]
            # res + identity_map (generator.py:44); after the last block also `x += head` (generator.py:161)
            ops.channelnorm(self.rows, self.g_flat if last else self.g_b1, blk.norm2.gamma, blk.norm2.beta,
                            act=ACT_NONE, reflect=not last, res1=x_f32, res2=self.head_f32 if last else None,
                            want_f32=not last, out_f32=nxt, out_act=self.act_flat if last else self.act_a)
            x_f32, x_act = nxt, (self.act_flat if last else self.act_a)
        if self.n_res == 0:
            raise RuntimeError("Generator: n_residual_blocks == 0 is not supported by the fused plan")
        h = x_act
        for i, (conv, g_cn, out_buf, cn_buf) in enumerate(self.ups):
            blk = getattr(mod, f"upconv_block{i + 1}")
            if g_cn is None:
                h = conv(h, blk[0].weight, blk[0].bias, blk[1].gamma, blk[1].beta, out=out_buf)
            else:
                rows = conv(h, blk[0].weight, blk[0].bias, out=out_buf)
                h, _ = ops.channelnorm(rows, g_cn, blk[1].gamma, blk[1].beta, act=ACT_RELU, out_act=cn_buf)
        last = mod.conv_block_out[1]
        return self.conv_out(h, last.weight, last.bias)


class HyperAnalysisPlan:
    def __init__(self, n, h, w, C, N, device):
        self.g_in = Geom(n, h, w, C, round_up(C, 64))
        self.in_act = self.g_in.alloc(device)
        b2 = (2, 2, 2, 2)
        g1 = Geom(n, h, w, N, round_up(N, 64), *b2)
        self.c1 = Conv(self.g_in, N, 3, pad_mode=PAD_ZERO, pad=(1, 1, 1, 1), out_geom=g1, out_reflect=True, act=ACT_RELU)
        h2, w2 = (h + 4 - 5) // 2 + 1, (w + 4 - 5) // 2 + 1
        g2 = Geom(n, h2, w2, N, round_up(N, 64), *b2)
        self.c2 = Conv(g1, N, 5, stride=2, pad_mode=PAD_REFLECT, pad=b2, out_geom=g2, out_reflect=True, act=ACT_RELU)
        self.c3 = Conv(g2, N, 5, stride=2, pad_mode=PAD_REFLECT, pad=b2, out_mode=OUT_NCHW_F32)
        self.b1, self.b2 = self.c1.alloc_out(device), self.c2.alloc_out(device)
        self.flops = self.c1.flops + self.c2.flops + self.c3.flops

    def run(self, mod, y):
        ops.nchw_to_act(y, self.g_in, out=self.in_act)
        h = self.c1(self.in_act, mod.conv1.weight, mod.conv1.bias, out=self.b1)
        h = self.c2(h, mod.conv2.weight, mod.conv2.bias, out=self.b2)
        return self.c3(h, mod.conv3.weight, mod.conv3.bias)


class HyperSynthesisPlan:
    def __init__(self, n, h, w, C, N, device):
        self.g_in = Geom(n, h, w, N, round_up(N, 64))
        self.in_act = self.g_in.alloc(device)
        g1 = Geom(n, 2 * h, 2 * w, N, round_up(N, 64))
        g2 = Geom(n, 4 * h, 4 * w, N, round_up(N, 64))
        self.c1 = Conv(self.g_in, N, 5, stride=2, transposed=True, pad=(2, 2, 2, 2), out_geom=g1, act=ACT_RELU)
        self.c2 = Conv(g1, N, 5, stride=2, transposed=True, pad=(2, 2, 2, 2), out_geom=g2, act=ACT_RELU)
        self.c3 = Conv(g2, C, 3, stride=1, transposed=True, pad=(1, 1, 1, 1), out_mode=OUT_NCHW_F32)
        self.b1, self.b2 = self.c1.alloc_out(device), self.c2.alloc_out(device)
        self.flops = self.c1.flops + self.c2.flops + self.c3.flops

    def run(self, mod, z):
        ops.nchw_to_act(z, self.g_in, out=self.in_act)
        h = self.c1(self.in_act, mod.conv1.weight, mod.conv1.bias, out=self.b1)
        h = self.c2(h, mod.conv2.weight, mod.conv2.bias, out=self.b2)
        return self.c3(h, mod.conv3.weight, mod.conv3.bias)


class HyperSynthesisDLMMPlan:
    """HyperpriorSynthesisDLMM.forward (src/network/hyper.py:121-130)."""

    def __init__(self, n, h, w, C, N, n_out, device):
        self.g_in = Geom(n, h, w, N, round_up(N, 64))
        self.in_act = self.g_in.alloc(device)
        g1 = Geom(n, 2 * h, 2 * w, N, round_up(N, 64))
        g2 = Geom(n, 4 * h, 4 * w, N, round_up(N, 64))
        g3 = Geom(n, 4 * h, 4 * w, C, round_up(C, 64))
        self.c1 = Conv(self.g_in, N, 5, stride=2, transposed=True, pad=(2, 2, 2, 2), out_geom=g1, act=ACT_RELU)
        self.c2 = Conv(g1, N, 5, stride=2, transposed=True, pad=(2, 2, 2, 2), out_geom=g2, act=ACT_RELU)
        self.c3 = Conv(g2, C, 3, stride=1, transposed=True, pad=(1, 1, 1, 1), out_geom=g3, act=ACT_NONE)
        self.c4 = Conv(g3, n_out, 1, out_mode=OUT_NCHW_F32)
        self.b1, self.b2, self.b3 = self.c1.alloc_out(device), self.c2.alloc_out(device), self.c3.alloc_out(device)
        self.flops = self.c1.flops + self.c2.flops + self.c3.flops + self.c4.flops

    def run(self, mod, z):
        ops.nchw_to_act(z, self.g_in, out=self.in_act)
        h = self.c1(self.in_act, mod.conv1.weight, mod.conv1.bias, out=self.b1)
        h = self.c2(h, mod.conv2.weight, mod.conv2.bias, out=self.b2)
        h = self.c3(h, mod.conv3.weight, mod.conv3.bias, out=self.b3)
        return self.c4(h, mod.conv_out.weight, mod.conv_out.bias)


class DiscriminatorPlan:
    """Discriminator.forward (src/network/discriminator.py:66-86) for n = 2B stacked real / generated images."""
    FILTERS = (64, 128, 256, 512)
    CONTEXT_C = 12

    def __init__(self, n, h, w, im_channels, C, ctx_h, ctx_w, device):
        if h % ctx_h or w % ctx_w or h // ctx_h != w // ctx_w:
            raise ValueError("Discriminator: image size must be an integer multiple of the context size")
        self.scale = h // ctx_h
        self.c_x = im_channels
        self.inv_sigmas = [None] * len(self.FILTERS)
        b1 = (1, 1, 1, 1)
        self.g_y = Geom(n, ctx_h, ctx_w, C, round_up(C, 64), *b1)
        self.y_act = self.g_y.alloc(device)
        self.g_ctx = Geom(n, ctx_h, ctx_w, self.CONTEXT_C, 16)
        self.c_ctx = Conv(self.g_y, self.CONTEXT_C, 3, pad_mode=PAD_REFLECT, pad=b1, out_geom=self.g_ctx, act=ACT_LEAKY02)
        self.ctx_act = self.c_ctx.alloc_out(device)
        self.g_in = Geom(n, h, w, im_channels + self.CONTEXT_C, 64, *b1)
        self.in_act = self.g_in.alloc(device)
        self.convs, self.bufs = [], []
        g = self.g_in
        for i, f in enumerate(self.FILTERS):
            last = i == len(self.FILTERS) - 1
            g_next = Geom(n, g.h // 2, g.w // 2, f, round_up(f, 64), *((0, 0, 0, 0) if last else b1))
            conv = Conv(g, f, 4, stride=2, pad_mode=PAD_REFLECT, pad=b1, out_geom=g_next, out_reflect=not last,
                        act=ACT_LEAKY02)
            self.convs.append(conv)
            self.bufs.append(conv.alloc_out(device))
            g = g_next
        self.c_out = Conv(g, 1, 1, out_mode=OUT_NCHW_F32)
        self.ws = [torch.empty(f + c.in_geom.c * 16, dtype=torch.float32, device=device)
                   for f, c in zip(self.FILTERS, self.convs)]
        self.flops = self.c_ctx.flops + sum(c.flops for c in self.convs) + self.c_out.flops

    def run(self, mod, x, y):
        # the training plan (train_plan.DiscriminatorTrainPlan) reads these buffers in its backward: any run() that is
        # not the one a pending backward belongs to bumps the generation so that backward fails loudly instead of
        # differentiating through the wrong activations
        self._generation = getattr(self, "_generation", 0) + 1
        ops.nchw_to_act(y, self.g_y, reflect=True, out=self.y_act)
        self.c_ctx(self.y_act, mod.context_conv.weight, mod.context_conv.bias, out=self.ctx_act)
        ops.disc_input(x, self.ctx_act, self.g_ctx, self.g_in, self.scale, out=self.in_act)
        h = self.in_act
        for i, conv in enumerate(self.convs):
            layer = getattr(mod, f"conv{i + 1}")
            # spectral norm (discriminator.py:46-62): one power iteration per training forward, W = W_orig / sigma
            _, inv_sigma = ops.spectral_sigma(layer.weight_orig, layer.weight_u, layer.weight_v, mod.training, self.ws[i])
            key = (layer.weight_u._version, layer.weight_v._version) if not mod.training else object()
            self.inv_sigmas[i] = inv_sigma
            h = conv(h, layer.weight_orig, layer.bias, out=self.bufs[i], scale=inv_sigma, scale_key=key)
        return self.c_out(h, mod.conv_out.weight, mod.conv_out.bias)


def wants_grad(module, *inputs):
    """True when the call must be recorded for autograd (training): then the module runs its training plan
    (hific_b200.train_plan) instead of the fused inference plan."""
    if not torch.is_grad_enabled():
        return False
    return any(t.requires_grad for t in inputs) or any(p.requires_grad for p in module.parameters())
