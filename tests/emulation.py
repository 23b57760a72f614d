"""TEST INFRASTRUCTURE: run the product's HOST glue of the compress path (shapes, coder layouts, padding, host rANS
calls, container) on a machine without a GPU by swapping every CUDA entry point it touches for the oracle's CPU
arithmetic.  Only tests import this; the product never does (it has no CPU path).  What this can and cannot show:
it pins the Python glue + the C host coder end to end against the reference's own `Model.compress` output
(tests/golden/entropy_coding.npz: model_m1 / model_m2); the CUDA kernels themselves are checked by the `-m gpu` tests.
"""
import contextlib

import numpy as np
import torch

from hific_b200 import engine, hyperprior, ops
from hific_b200._lib import SYM_PIXEL_STEPS
from hific_b200.compression import hyperprior_model
from hific_b200.network import encoder, generator, hyper
from oracle import entropy_oracle as EO
from oracle import hific_oracle as O


def _to_layout(t, layout):
    if layout == SYM_PIXEL_STEPS:
        return t.permute(0, 2, 3, 1).reshape(-1).contiguous()
    return t.reshape(-1).contiguous()


def _from_layout(flat, shape, layout):
    n, c, h, w = shape
    if layout == SYM_PIXEL_STEPS:
        return flat.reshape(n, h, w, c).permute(0, 3, 1, 2).contiguous()
    return flat.reshape(n, c, h, w)


def quantize_symbols(x, mean=None, scale_raw=None, scale_table=None, scale_lower_bound=0.11, likelihood_type="gaussian",
                     layout=0, want_symbols=True, want_indices=True, want_dequant=False, want_bits=False):
    n, c, h, w = x.shape
    out = {}
    sym = torch.floor(x + 0.5 - mean) if mean is not None else torch.floor(x + 0.5)
    sym = sym.to(torch.int32)
    if want_symbols:
        out["symbols"] = _to_layout(sym, layout)
    if want_indices:
        if scale_raw is not None:
            idx = EO.compute_indices(torch.clamp(scale_raw, scale_lower_bound), scale_table)
        else:
            idx = torch.from_numpy(EO.hyper_indices((n, c, h, w)))
        out["indices"] = _to_layout(idx, layout)
    if want_dequant:
        out["dequant"] = sym.float() + mean if mean is not None else sym.float()
    if want_bits:
        bits = EO.prior_bits(x, mean, torch.clamp(scale_raw, scale_lower_bound), likelihood_type)
        out["bits_sum"] = (bits * -np.log(2.)).to(torch.float64)
    return out


def scale_indices(scale_raw, scale_table, scale_lower_bound=0.11, layout=0):
    return _to_layout(EO.compute_indices(torch.clamp(scale_raw, scale_lower_bound), scale_table), layout)


def dequantize_symbols(symbols, mean, shape, layout=0):
    s = _from_layout(symbols, shape, layout).float()
    return s + mean if mean is not None else s


def _sd(module):
    return {k: v.detach() for k, v in module.state_dict().items()}


@contextlib.contextmanager
def cpu_emulation():
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    patch(ops, "quantize_symbols", quantize_symbols)
    patch(ops, "scale_indices", scale_indices)
    patch(ops, "dequantize_symbols", dequantize_symbols)
    patch(engine, "_require_cuda", lambda x, who: None)
    patch(encoder.Encoder, "forward", lambda self, x: O.encoder_forward(_sd(self), x, prefix=""))
    patch(generator.Generator, "forward", lambda self, y: O.generator_forward(_sd(self), y, n_residual_blocks=self.n_residual_blocks, prefix=""))
    patch(hyper.HyperpriorAnalysis, "forward", lambda self, y: O.hyper_analysis(_sd(self), y, prefix=""))
    patch(hyper.HyperpriorSynthesis, "forward", lambda self, z: O.hyper_synthesis(_sd(self), z, prefix=""))
    patch(hyperprior.Hyperprior, "_latent_statistics",
          lambda self, z: (self.synthesis_mu(z).contiguous(), self.synthesis_std(z).contiguous()))

    def hyper_bits(self, x, spatial_shape):
        params = {k: v.detach() for k, v in self.distribution.state_dict().items()}
        n_bits = EO.hyper_bits(x, params).to(torch.float32)
        return n_bits, n_bits / float(np.prod(spatial_shape)), n_bits / x.shape[0]

    patch(hyperprior_model.HyperpriorEntropyModel, "_estimate_compression_bits", hyper_bits)
    try:
        yield
    finally:
        for obj, name, value in reversed(saved):
            setattr(obj, name, value)


# ----------------------------------------------------------------------------------------------------------------------
# LPIPS trunk plan (hific_b200.loss.lpips_trunk): CPU stand-ins for the entry points it calls, written independently of
# the CUDA code (plain torch ops on NCHW tensors), so that the PLAN (geometry, weight re-indexing, batching of the two
# images, order of the adjoints) can be checked against torchvision's AlexNet + autograd without a GPU.  The `-m gpu`
# tests then compare the real kernels with these same functions.
# ----------------------------------------------------------------------------------------------------------------------
import torch.nn.functional as F

from hific_b200 import grad as _grad
from hific_b200._lib import ACT_RELU, OUT_NHWC_F16, PAD_ZERO


def _act_to_nchw(x_act, geom):
    return x_act.view(geom.shape)[..., :geom.c].permute(0, 3, 1, 2).float()


def _nchw_to_act(t, geom, out):
    out.view(geom.shape).zero_()
    out.view(geom.shape)[..., :geom.c] = t.permute(0, 2, 3, 1).to(torch.float16)
    return out


def conv_call(self, x_act, weight, bias=None, gamma=None, beta=None, out=None, scale=None, scale_key=None):
    """ops.Conv.__call__ for the configuration the LPIPS plan uses: stride-1 conv2d, zero padding, bias + ReLU,
    border-less NHWC fp16 in and out, fp16 operands with fp32 accumulation."""
    d = self.desc
    if d.stride != 1 or d.transposed or d.pad_mode != PAD_ZERO or d.out_mode != OUT_NHWC_F16 or d.norm or d.window \
            or d.act != ACT_RELU or any((self.in_geom.pt, self.in_geom.pl, self.out_geom.pt, self.out_geom.pl)):
        raise NotImplementedError("emulation covers the LPIPS trunk's conv configuration only")
    x = _act_to_nchw(x_act, self.in_geom)
    w = weight.detach().to(torch.float16).float()
    y = F.conv2d(F.pad(x, (d.pad_l, d.pad_r, d.pad_t, d.pad_b)), w, bias.detach().float() if bias is not None else None)
    y = torch.relu(y)
    if out is None:
        out = self.out_geom.alloc(x_act.device)
    return _nchw_to_act(y, self.out_geom, out)


def convgrad_data_grad(self, dy_rows, weight, out=None, scale=None, dy_act=None):
    """grad.ConvGrad.data_grad for a stride-1 zero-padded conv2d: bf16 operands, fp32 accumulation, fp32 rows out."""
    if self.transposed or self.stride != 1 or self.pad_mode != PAD_ZERO:
        raise NotImplementedError
    n = self.in_geom.n
    dy = dy_rows[:, :self.cout].reshape(n, self.oh, self.ow, self.cout).permute(0, 3, 1, 2)
    dy = dy.to(_GFMT).float()
    w = weight.detach().to(_GFMT).float()
    dx = F.conv_transpose2d(dy, w, stride=1, padding=self.pad[0])
    rows = torch.zeros((n * self.in_geom.h * self.in_geom.w, self.cin4), dtype=torch.float32)
    rows[:, :self.cin] = dx.permute(0, 2, 3, 1).reshape(-1, self.cin)
    return rows


def lpips_prep(target, pred, geom, normalize, shift, scale, out=None):
    x = torch.cat([target, pred], 0).float()
    if normalize:
        x = 2 * x - 1
    x = (x - shift.view(1, 3, 1, 1)) / scale.view(1, 3, 1, 1)
    n2, _, h, w = x.shape
    xp = torch.zeros((n2, 3, 4 * geom.h, 4 * geom.w))
    hh, ww = min(h, 4 * geom.h - 2), min(w, 4 * geom.w - 2)
    xp[:, :, 2:2 + hh, 2:2 + ww] = x[:, :, :hh, :ww]
    s2d = xp.view(n2, 3, geom.h, 4, geom.w, 4).permute(0, 3, 5, 1, 2, 4).reshape(n2, 48, geom.h, geom.w)
    if out is None:
        out = geom.alloc(x.device)
    return _nchw_to_act(s2d, geom, out)


def lpips_prep_bwd(g_rows, n, h, w, hs, ws, normalize, scale):
    g = g_rows[:, :48].reshape(n, hs, ws, 4, 4, 3).permute(0, 5, 1, 3, 2, 4).reshape(n, 3, 4 * hs, 4 * ws)
    d = torch.zeros((n, 3, h, w))
    hh, ww = min(h, 4 * hs - 2), min(w, 4 * ws - 2)
    d[:, :, :hh, :ww] = g[:, :, 2:2 + hh, 2:2 + ww]
    return d * (2.0 if normalize else 1.0) / scale.view(1, 3, 1, 1)


def maxpool3s2(x_act, geom, out_geom, out=None):
    y = F.max_pool2d(_act_to_nchw(x_act, geom), 3, 2)
    if out is None:
        out = out_geom.alloc(x_act.device)
    return _nchw_to_act(y, out_geom, out)


def maxpool3s2_bwd(g_out_rows, x_act, geom):
    with torch.enable_grad():                 # called from inside an autograd Function's backward (grad mode off)
        x = _act_to_nchw(x_act, geom).requires_grad_(True)
        y = F.max_pool2d(x, 3, 2)
        g = g_out_rows[:, :geom.c].reshape(geom.n, y.shape[2], y.shape[3], geom.c).permute(0, 3, 1, 2)
        y.backward(g)
    return x.grad.permute(0, 2, 3, 1).reshape(-1, geom.c).contiguous()


def _lpips_dist(f, lin_w, n):
    f0, f1 = f[:n], f[n:]
    n0 = f0 / torch.sqrt((f0 ** 2).sum(1, keepdim=True) + 1e-10)
    n1 = f1 / torch.sqrt((f1 ** 2).sum(1, keepdim=True) + 1e-10)
    return (((n0 - n1) ** 2) * lin_w.view(1, -1, 1, 1)).sum(1).mean(dim=(1, 2))


def lpips_nhwc(feat_act, geom, lin_w, out):
    out += _lpips_dist(_act_to_nchw(feat_act, geom), lin_w.detach().float(), geom.n // 2)
    return out


def lpips_nhwc_bwd(feat_act, geom, lin_w, upstream, g_in_rows=None):
    n = geom.n // 2
    f = _act_to_nchw(feat_act, geom)
    with torch.enable_grad():
        f1 = f[n:].clone().requires_grad_(True)
        d = _lpips_dist(torch.cat([f[:n], f1], 0), lin_w.detach().float(), n)
        d.backward(upstream.float())
    g = f1.grad
    if g_in_rows is not None:
        g = g + g_in_rows[:, :geom.c].reshape(n, geom.h, geom.w, geom.c).permute(0, 3, 1, 2)
    g = g * (f[n:] > 0)
    return g.permute(0, 2, 3, 1).reshape(-1, geom.c).contiguous()


@contextlib.contextmanager
def lpips_cpu_emulation():
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    patch(ops.Conv, "__call__", conv_call)
    patch(_grad.ConvGrad, "data_grad", convgrad_data_grad)
    for name, fn in (("lpips_prep", lpips_prep), ("lpips_prep_bwd", lpips_prep_bwd), ("maxpool3s2", maxpool3s2),
                     ("maxpool3s2_bwd", maxpool3s2_bwd), ("lpips_nhwc", lpips_nhwc), ("lpips_nhwc_bwd", lpips_nhwc_bwd)):
        patch(ops, name, fn)
    try:
        yield
    finally:
        for obj, name, value in reversed(saved):
            setattr(obj, name, value)


# ----------------------------------------------------------------------------------------------------------------------
# General CPU stand-ins for the activation-format entry points the inference plans (hific_b200.engine) call, so that
# the plans' geometry -- materialised reflect borders, asymmetric pads, stride-2 / transposed layers, channel padding,
# fused ChannelNorm, residual adds, row pitches -- can be checked against the oracle without a GPU.  They model the
# numerics of the kernels only to first order (operands rounded to fp16, fp32 accumulation in torch's order).
# ----------------------------------------------------------------------------------------------------------------------
from hific_b200._lib import ACT_LEAKY02, ACT_NONE, OUT_NCHW_F32, OUT_NHWC_F32, PAD_REFLECT

CN_EPS = 1e-3


def _apply_act(y, act):
    if act == ACT_RELU:
        return torch.relu(y)
    if act == ACT_LEAKY02:
        return F.leaky_relu(y, 0.2)
    return y


def _channel_norm(y, gamma, beta):
    mean = y.mean(dim=1, keepdim=True)
    var = y.var(dim=1, keepdim=True)                       # unbiased, as torch.var in channel.py:48-59
    return gamma.view(1, -1, 1, 1) * (y - mean) * torch.rsqrt(var + CN_EPS) + beta.view(1, -1, 1, 1)


def _write_act(y, geom, out, reflect):
    """(n, c, h, w) fp32 -> bordered NHWC fp16 buffer; the border by reflection of the interior (ReflectionPad2d
    semantics) when `reflect`, otherwise left at zero."""
    buf = out.view(geom.shape)
    buf.zero_()
    t = y
    if reflect and any((geom.pt, geom.pl, geom.pb, geom.pr)):
        t = F.pad(y, (geom.pl, geom.pr, geom.pt, geom.pb), mode="reflect")
        buf[..., :geom.c] = t.permute(0, 2, 3, 1).to(torch.float16)
    else:
        buf[:, geom.pt:geom.pt + geom.h, geom.pl:geom.pl + geom.w, :geom.c] = t.permute(0, 2, 3, 1).to(torch.float16)
    return out


def conv_call_widenorm(self, x_act, weight, bias, gamma, beta, res1=None, res2=None, out_f32=None, out_act=None):
    """ops.Conv.call_widenorm: conv + ChannelNorm over the whole channel row + act [+ res1] [+ res2]."""
    go = self.out_geom
    y = _conv_linear(self, x_act, weight, bias)
    y = _apply_act(_channel_norm(y, gamma.detach().float().reshape(-1), beta.detach().float().reshape(-1)), self.desc.act)
    for r in (res1, res2):
        if r is not None:
            y = y + r.view(go.n, go.h, go.w, -1)[..., :self.cout].permute(0, 3, 1, 2)
    if out_f32 is not None:
        out_f32.view(go.n * go.h * go.w, -1)[:, :self.cout] = y.permute(0, 2, 3, 1).reshape(-1, self.cout)
    if out_act is not None:
        _write_act(y, go, out_act, bool(self.desc.out_reflect))
    return out_act, out_f32


def _conv_linear(self, x_act, weight, bias=None, scale=None):
    d, gi = self.desc, self.in_geom
    buf = x_act.view(gi.shape).float()
    w = weight.detach().float()
    if scale is not None:
        w = w * scale.float().reshape(())
    w = (w.to(torch.bfloat16) if d.b_bf16 else w.to(torch.float16)).float()
    if d.dgrad:
        w = w.flip(2, 3).transpose(0, 1)
    interior = buf[:, gi.pt:gi.pt + gi.h, gi.pl:gi.pl + gi.w, :gi.c].permute(0, 3, 1, 2)
    if d.transposed:
        y = F.conv_transpose2d(interior, w, stride=d.stride, padding=d.pad_t, output_padding=d.stride - 1)
    else:
        if d.pad_mode == PAD_REFLECT:
            region = buf[:, gi.pt - d.pad_t:gi.pt + gi.h + d.pad_b, gi.pl - d.pad_l:gi.pl + gi.w + d.pad_r, :gi.c]
            region = region.permute(0, 3, 1, 2)
        else:
            region = F.pad(interior, (d.pad_l, d.pad_r, d.pad_t, d.pad_b))
        y = F.conv2d(region, w, stride=d.stride)
    if bias is not None:
        y = y + bias.detach().float().view(1, -1, 1, 1)
    return y


def conv_call_general(self, x_act, weight, bias=None, gamma=None, beta=None, out=None, scale=None, scale_key=None):
    d, gi = self.desc, self.in_geom
    buf = x_act.view(gi.shape).float()
    w = weight.detach().float()
    if scale is not None:
        w = w * scale.float().reshape(())
    w = (w.to(torch.bfloat16) if d.b_bf16 else w.to(torch.float16)).float()
    if d.dgrad:
        w = w.flip(2, 3).transpose(0, 1)
    interior = buf[:, gi.pt:gi.pt + gi.h, gi.pl:gi.pl + gi.w, :gi.c].permute(0, 3, 1, 2)
    if d.transposed:
        assert not any((gi.pt, gi.pl, gi.pb, gi.pr))
        y = F.conv_transpose2d(interior, w, stride=d.stride, padding=d.pad_t, output_padding=d.stride - 1)
    else:
        if d.pad_mode == PAD_REFLECT:
            assert gi.pt >= d.pad_t and gi.pl >= d.pad_l and gi.pb >= d.pad_b and gi.pr >= d.pad_r
            region = buf[:, gi.pt - d.pad_t:gi.pt + gi.h + d.pad_b, gi.pl - d.pad_l:gi.pl + gi.w + d.pad_r, :gi.c]
            region = region.permute(0, 3, 1, 2)             # reads the MATERIALISED border the producer wrote
        else:
            region = F.pad(interior, (d.pad_l, d.pad_r, d.pad_t, d.pad_b))
        y = F.conv2d(region, w, stride=d.stride)
    if bias is not None:
        y = y + bias.detach().float().view(1, -1, 1, 1)
    if d.norm:
        y = _channel_norm(y, gamma.detach().float().reshape(-1), beta.detach().float().reshape(-1))
    y = _apply_act(y, d.act)
    go = self.out_geom
    assert tuple(y.shape[2:]) == (go.h, go.w), (y.shape, go)
    if out is None:
        out = self.alloc_out(x_act.device)
    if d.out_mode == OUT_NHWC_F16:
        return _write_act(y, go, out, bool(d.out_reflect))
    if d.out_mode == OUT_NHWC_F32:
        out.zero_()
        out.view(-1, go.cpad)[:, :self.cout] = y.permute(0, 2, 3, 1).reshape(-1, self.cout)
        return out
    out.copy_(y)
    return out


def nchw_to_act(x, geom, reflect=False, norm=False, gamma=None, beta=None, out=None):
    y = x.float()
    if norm:
        y = _channel_norm(y, gamma.detach().float().reshape(-1), beta.detach().float().reshape(-1))
    if out is None:
        out = geom.alloc(x.device)
    return _write_act(y, geom, out, reflect)


def channelnorm(x_rows, geom, gamma, beta, act=ACT_NONE, reflect=False, res1=None, res2=None, want_f32=False,
                want_act=True, out_act=None, out_f32=None):
    n, h, w, c = geom.n, geom.h, geom.w, geom.c
    y = x_rows.view(n * h * w, -1)[:, :c].reshape(n, h, w, c).permute(0, 3, 1, 2)
    y = _apply_act(_channel_norm(y, gamma.detach().float().reshape(-1), beta.detach().float().reshape(-1)), act)
    for r in (res1, res2):
        if r is not None:
            y = y + r.view(n, h, w, -1)[..., :c].permute(0, 3, 1, 2)
    if want_f32:
        if out_f32 is None:
            out_f32 = torch.empty((n * h * w, c), dtype=torch.float32)
        out_f32.copy_(y.permute(0, 2, 3, 1).reshape(n * h * w, c))
    if want_act:
        if out_act is None:
            out_act = geom.alloc(x_rows.device)
        _write_act(y, geom, out_act, reflect)
    return out_act, out_f32


def instancenorm(x_rows, geom, gamma, beta, act=ACT_NONE, reflect=False, res1=None, res2=None, want_f32=False,
                 want_act=True, out_act=None, out_f32=None):
    """ops.instancenorm: torch.nn.functional.instance_norm on the NCHW view of the rows (eps 1e-5, biased variance)."""
    n, h, w, c = geom.n, geom.h, geom.w, geom.c
    y = x_rows.view(n * h * w, -1)[:, :c].reshape(n, h, w, c).permute(0, 3, 1, 2)
    y = F.instance_norm(y, weight=gamma.detach().float().reshape(-1), bias=beta.detach().float().reshape(-1), eps=ops.IN_EPS)
    y = _apply_act(y, act)
    for r in (res1, res2):
        if r is not None:
            y = y + r.view(n, h, w, -1)[..., :c].permute(0, 3, 1, 2)
    if want_f32:
        if out_f32 is None:
            out_f32 = torch.empty((n * h * w, c), dtype=torch.float32)
        out_f32.copy_(y.permute(0, 2, 3, 1).reshape(n * h * w, c))
    if want_act:
        if out_act is None:
            out_act = geom.alloc(x_rows.device)
        _write_act(y, geom, out_act, reflect)
    return out_act, out_f32


@contextlib.contextmanager
def plan_cpu_emulation():
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    patch(ops.Conv, "__call__", conv_call_general)
    patch(ops.Conv, "call_widenorm", conv_call_widenorm)
    patch(ops, "nchw_to_act", nchw_to_act)
    patch(ops, "channelnorm", channelnorm)
    patch(ops, "instancenorm", instancenorm)
    patch(engine, "_require_cuda", lambda x, who: None)
    try:
        yield
    finally:
        for obj, name, value in reversed(saved):
            setattr(obj, name, value)


# ----------------------------------------------------------------------------------------------------------------------
# Whole `Model.forward` on the CPU: the plan stand-ins above + the two likelihood entry points + inert CUDA streams.
# ----------------------------------------------------------------------------------------------------------------------
def latent_likelihood(y, mean, scale_raw, noise=None, scale_lower_bound=0.11, likelihood_type="gaussian", sums=None):
    scale = torch.clamp(scale_raw, min=scale_lower_bound)
    if sums is None:
        sums = torch.zeros(2, dtype=torch.float64)
    qy = torch.floor(y - mean + 0.5) + mean
    sums[1] += torch.log(O.latent_likelihood(qy, mean, scale, likelihood_type) + 1e-9).double().sum()
    if noise is not None:
        sums[0] += torch.log(O.latent_likelihood(y + noise, mean, scale, likelihood_type) + 1e-9).double().sum()
    return O.quantize_st(y, mean), sums


def _density_logits_packed(x, p):
    """x (n, c, hw), p (c, 64) packed as documented in include/hfc.h (softplus / tanh already applied)."""
    c = p.shape[0]
    P = lambda lo, hi, *shape: p[:, lo:hi].reshape(c, *shape)
    h = x.permute(1, 0, 2).reshape(c, 1, -1)                               # (c, 1, N)
    layers = ((P(0, 3, 3, 1), P(3, 6, 3, 1), P(6, 9, 3, 1)), (P(9, 18, 3, 3), P(18, 21, 3, 1), P(21, 24, 3, 1)),
              (P(24, 33, 3, 3), P(33, 36, 3, 1), P(36, 39, 3, 1)), (P(39, 42, 1, 3), P(42, 43, 1, 1), P(43, 44, 1, 1)))
    for H, b, a in layers:
        h = torch.bmm(H, h) + b
        h = h + a * torch.tanh(h)
    return h.reshape(c, x.shape[0], -1).permute(1, 0, 2)


def hyperlatent_likelihood(z, params64, noise=None, sums=None):
    n, c, hh, ww = z.shape
    if sums is None:
        sums = torch.zeros(2, dtype=torch.float64)

    def loglik(v):
        x = v.reshape(n, c, -1)
        up, lo = _density_logits_packed(x + 0.5, params64), _density_logits_packed(x - 0.5, params64)
        sign = -torch.sign(up + lo)
        p = torch.clamp(torch.abs(torch.sigmoid(sign * up) - torch.sigmoid(sign * lo)), min=1e-9)
        return torch.log(p + 1e-9).double().sum()

    z_quant = torch.floor(z + 0.5)
    sums[1] += loglik(z_quant)
    z_noisy = None
    if noise is not None:
        z_noisy = z + noise
        sums[0] += loglik(z_noisy)
    return z_noisy, z_quant, sums


class _InertStream:
    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass


@contextlib.contextmanager
def model_cpu_emulation():
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    with plan_cpu_emulation():
        patch(ops, "latent_likelihood", latent_likelihood)
        patch(ops, "hyperlatent_likelihood", hyperlatent_likelihood)
        patch(torch.cuda, "current_stream", lambda *a, **k: _InertStream())
        patch(torch.cuda, "stream", lambda s: contextlib.nullcontext())
        patch(hyperprior.Hyperprior, "_side_stream", lambda self, device: _InertStream())
        patch(torch.Tensor, "record_stream", lambda self, s: None)
        try:
            yield
        finally:
            for obj, name, value in reversed(saved):
                setattr(obj, name, value)


# ----------------------------------------------------------------------------------------------------------------------
# Discretised mixture likelihood (csrc/dlmm.cu): the forward through the oracle, the backward as the SAME closed-form
# expressions the kernel evaluates (so that the formulas themselves are checked against torch autograd on the CPU).
# ----------------------------------------------------------------------------------------------------------------------
def dlmm_likelihood(x, dlmm_params, noise=None, likelihood_type="gaussian", straight_through=True, sums=None):
    if sums is None:
        sums = torch.zeros(2, dtype=torch.float64)
    q = torch.floor(x + 0.5)
    sums[1] += O.dlmm_log_likelihood(q, dlmm_params, likelihood_type).double().sum()
    if noise is not None:
        sums[0] += O.dlmm_log_likelihood(x + noise, dlmm_params, likelihood_type).double().sum()
    return (x + (q - x)) if straight_through else q, sums


def dlmm_likelihood_bwd(x, dlmm_params, noise, d_decoded, g_nbpp, likelihood_type="gaussian"):
    n, c, h, w = x.shape
    k = dlmm_params.shape[1] // (3 * c)
    p = dlmm_params.reshape(n, 3, c, k, h, w)
    logit, mu, ls_raw = p[:, 0], p[:, 1], p[:, 2]
    g = float(g_nbpp.reshape(-1)[0])
    cdf = (lambda t: 0.5 * torch.erfc(t * (-1.0 / 2 ** 0.5))) if likelihood_type == "gaussian" else torch.sigmoid
    if likelihood_type == "gaussian":
        pdf = lambda t: 0.39894228040143267794 * torch.exp(-0.5 * t * t)
    else:
        pdf = lambda t: torch.sigmoid(t) * (1 - torch.sigmoid(t))
    v = (x + noise).reshape(n, c, 1, h, w)
    inv = torch.exp(-torch.clamp(ls_raw, min=-3.0))
    d = torch.abs(v - mu)
    tu, tl = inv * (0.5 - d), inv * (-0.5 - d)
    praw = cdf(tu) - cdf(tl)
    lse = torch.logsumexp(logit, dim=2, keepdim=True)
    a = (logit - lse) + torch.log(torch.clamp(praw, min=1e-9))
    wgt = torch.softmax(a, dim=2)
    pi = torch.exp(logit - lse)
    d_logit = g * (wgt - pi)
    gp = g * wgt / torch.clamp(praw, min=1e-9)
    gp = torch.where((praw >= 1e-9) | (gp < 0), gp, torch.zeros_like(gp))
    fu, fl = pdf(tu), pdf(tl)
    dp_dd = -inv * (fu - fl)
    dp_dinv = fu * (0.5 - d) - fl * (-0.5 - d)
    sgn = torch.sign(v - mu)
    dv = (gp * dp_dd * sgn).sum(dim=2)
    d_mu = -gp * dp_dd * sgn
    gls = gp * dp_dinv * -inv
    gls = torch.where((ls_raw >= -3.0) | (gls < 0), gls, torch.zeros_like(gls))
    dparams = torch.stack([d_logit, d_mu, gls], dim=1).reshape(dlmm_params.shape)
    dx = dv + (d_decoded if d_decoded is not None else 0)
    return dx, dparams


# ----------------------------------------------------------------------------------------------------------------------
# Training path on the CPU: grad.ConvGrad through torch autograd of the same convolution (bf16-rounded operands, as the
# backward GEMMs use), train_plan.relu_mask, and a differentiable stand-in of the hyper-latent likelihood Function.
# ----------------------------------------------------------------------------------------------------------------------
from hific_b200 import train_plan as _train_plan


_GFMT = torch.bfloat16 if _grad.GRAD_BF16 else torch.float16     # gradient operand format of the product (grad.py)


def _convgrad_apply(cg, x, w):
    if cg.transposed:
        return F.conv_transpose2d(x, w, stride=cg.stride, padding=cg.pad[0], output_padding=cg.stride - 1)
    pt, pl, pb, pr = cg.pad
    xp = F.pad(x, (pl, pr, pt, pb), mode="reflect" if cg.pad_mode == PAD_REFLECT else "constant")
    return F.conv2d(xp, w, stride=cg.stride)


def _rows_to_nchw(rows, n, h, w, c):
    return rows[:, :c].reshape(n, h, w, c).permute(0, 3, 1, 2)


def convgrad_data_grad_general(self, dy_rows, weight, out=None, scale=None, dy_act=None):
    assert dy_rows is not None, "emulation needs the fp32 gradient rows"
    g = self.in_geom
    dy = _rows_to_nchw(dy_rows, g.n, self.oh, self.ow, self.cout).to(_GFMT).float()
    w = weight.detach().float()
    if scale is not None:
        w = w * scale.float().reshape(())
    w = w.to(_GFMT).float()
    with torch.enable_grad():
        x = torch.zeros((g.n, self.cin, g.h, g.w), requires_grad=True)
        _convgrad_apply(self, x, w).backward(dy)
    rows = torch.zeros((g.n * g.h * g.w, self.cin4))
    rows[:, :self.cin] = x.grad.permute(0, 2, 3, 1).reshape(-1, self.cin)
    return rows


def convgrad_weight_grad(self, x_act, dy_rows, dw_out=None, accumulate=False, scale=1.0, dy_act=None):
    assert dy_rows is not None, "emulation needs the fp32 gradient rows"
    g = self.in_geom
    x = x_act.view(g.shape)[:, g.pt:g.pt + g.h, g.pl:g.pl + g.w, :g.c].permute(0, 3, 1, 2).float()
    x = x.to(_GFMT).float()
    dy = _rows_to_nchw(dy_rows, g.n, self.oh, self.ow, self.cout).to(_GFMT).float()
    shape = (self.cin, self.cout, self.k, self.k) if self.transposed else (self.cout, self.cin, self.k, self.k)
    with torch.enable_grad():
        w = torch.zeros(shape, requires_grad=True)
        _convgrad_apply(self, x, w).backward(dy)
    dw = w.grad * (scale * _grad._inv_scale)
    if dw_out is None:
        return dw
    if accumulate:
        dw_out += dw
    else:
        dw_out.copy_(dw)
    return dw_out


def convgrad_bias_grad(self, dy_rows, db_out=None, accumulate=False, scale=1.0):
    db = dy_rows[:, :self.cout].sum(0) * (scale * _grad._inv_scale)
    if db_out is None:
        return db
    if accumulate:
        db_out += db
    else:
        db_out.copy_(db)
    return db_out


def relu_mask(g, y_act, geom, slope=0.0):
    y = y_act.view(geom.shape)[:, geom.pt:geom.pt + geom.h, geom.pl:geom.pl + geom.w, :geom.c].reshape(-1, geom.c).float()
    out = torch.zeros((g.shape[0], (geom.c + 3) // 4 * 4))
    out[:, :geom.c] = g[:, :geom.c] * torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope))
    return out


def hyperlatent_likelihood_fn(z, params64, noise):
    """Differentiable stand-in of ops.HyperlatentLikelihoodFn.apply."""
    n, c = z.shape[:2]

    def loglik(v):
        x = v.reshape(n, c, -1)
        up, lo = _density_logits_packed(x + 0.5, params64), _density_logits_packed(x - 0.5, params64)
        sign = -torch.sign(up + lo).detach()
        p = O.lower_bound(torch.abs(torch.sigmoid(sign * up) - torch.sigmoid(sign * lo)), 1e-9)
        return torch.log(p + 1e-9).double().sum()

    z_noisy, z_quant = z + noise, torch.floor(z + 0.5).detach()
    return z_noisy, z_quant, torch.stack([loglik(z_noisy), loglik(z_quant).detach()])


@contextlib.contextmanager
def training_cpu_emulation():
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    with model_cpu_emulation():
        patch(_grad.ConvGrad, "data_grad", convgrad_data_grad_general)
        patch(_grad.ConvGrad, "weight_grad", convgrad_weight_grad)
        patch(_grad.ConvGrad, "bias_grad", convgrad_bias_grad)
        patch(_grad.ConvGrad, "dy_to_act", lambda self, dy_rows: None)     # the stand-ins work from the fp32 rows
        patch(_train_plan, "relu_mask", relu_mask)
        patch(_train_plan, "norm_bwd", norm_bwd)
        patch(ops, "dlmm_likelihood", dlmm_likelihood)
        patch(ops, "dlmm_likelihood_bwd", dlmm_likelihood_bwd)
        patch(ops.HyperlatentLikelihoodFn, "apply", staticmethod(hyperlatent_likelihood_fn))
        try:
            yield
        finally:
            for obj, name, value in reversed(saved):
                setattr(obj, name, value)


def norm_bwd(z, g, gamma, beta, act, as_operand=True, kind="channel", n=None):
    """train_plan.norm_bwd through torch autograd of ChannelNorm / InstanceNorm (+ReLU) on fp32 rows; always returns fp32
    rows (the emulated ConvGrad works from rows, so the 16-bit operand form of the real kernels is not modelled)."""
    c = gamma.numel()
    with torch.enable_grad():
        zz = z[:, :c].detach().clone().requires_grad_(True)
        gm = gamma.detach().reshape(-1).clone().requires_grad_(True)
        bt = beta.detach().reshape(-1).clone().requires_grad_(True)
        if kind == "instance":
            img = zz.view(n, -1, c).permute(0, 2, 1)                     # (n, c, hw)
            y = F.instance_norm(img, weight=gm, bias=bt, eps=ops.IN_EPS).permute(0, 2, 1).reshape(-1, c)
            y = _apply_act(y, act)
        else:
            mean = zz.mean(dim=1, keepdim=True)
            var = zz.var(dim=1, keepdim=True)
            y = _apply_act(gm * (zz - mean) * torch.rsqrt(var + CN_EPS) + bt, act)
        y.backward(g[:, :c])
    dz = torch.zeros((z.shape[0], (c + 3) // 4 * 4))
    dz[:, :c] = zz.grad
    inv = _grad._inv_scale
    return dz, (gm.grad * inv).view_as(gamma), (bt.grad * inv).view_as(beta), zz.grad.sum(0) * inv


# ----------------------------------------------------------------------------------------------------------------------
# Differentiable stand-ins of the remaining autograd Functions of the compression model's training step, so that
# `Model.compression_forward` + losses + `backward()` run end to end on the CPU (host logic of the whole step).
# ----------------------------------------------------------------------------------------------------------------------
def latent_likelihood_fn(y, mean, scale_raw, noise, lb, kind):
    """ops.LatentLikelihoodFn.apply: (decoded, sums[2]) with LowerBoundToward gates and straight-through latents."""
    scale = O.lower_bound(scale_raw, lb)
    qy = torch.floor(y - mean + 0.5).detach() + mean.detach()
    s_n = torch.log(O.latent_likelihood(y + noise, mean, scale, kind) + 1e-9).double().sum()
    s_q = torch.log(O.latent_likelihood(qy, mean.detach(), scale.detach(), kind) + 1e-9).double().sum().detach()
    # quantize_latents_st (hyperprior.py:108-122): d decoded / d y = 1, the mean carries no gradient
    decoded = O.quantize_st(y, mean.detach())
    return decoded, torch.stack([s_n, s_q])


def sqdiff_mean_fn(a, b, scale):
    return torch.mean((a * scale - b * scale) ** 2)


@contextlib.contextmanager
def train_step_cpu_emulation():
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    with training_cpu_emulation():
        patch(ops.LatentLikelihoodFn, "apply", staticmethod(latent_likelihood_fn))
        patch(ops.SqDiffMeanFn, "apply", staticmethod(sqdiff_mean_fn))
        try:
            yield
        finally:
            for obj, name, value in reversed(saved):
                setattr(obj, name, value)


# ----------------------------------------------------------------------------------------------------------------------
# Drop-in test (tests/test_dropin_reference.py): the reference's train.py / src/model.py drive the mirror's networks on the
# CPU.  Encoder / Generator / hyper networks / Hyperprior run their REAL host code (training plans, autograd Functions)
# over the stand-ins above; the Discriminator's plan has no CPU stand-ins for its layout / spectral-norm entry points, so
# its forward is the oracle's with torch autograd -- including torch.nn.utils.spectral_norm's in-place u / v update.
# ----------------------------------------------------------------------------------------------------------------------
from hific_b200.network import discriminator as _discriminator


def discriminator_forward(self, x, y):
    """Discriminator.forward (src/network/discriminator.py:66-86) on fp16-rounded conv operands."""
    if x.shape[0] != y.shape[0]:
        raise ValueError("Discriminator: image and context batch sizes differ")
    sd = dict(self.named_parameters())
    sd.update(dict(self.named_buffers()))
    out, logits, new_uv = O.discriminator_forward(sd, x, y, prefix="", training=self.training, rnd=O.round_fp16)
    if self.training:
        with torch.no_grad():
            for name, (u, v) in new_uv.items():
                getattr(self, name).weight_u.copy_(u)
                getattr(self, name).weight_v.copy_(v)
    return out, logits


@contextlib.contextmanager
def dropin_cpu_emulation():
    with train_step_cpu_emulation():
        old = _discriminator.Discriminator.forward
        _discriminator.Discriminator.forward = discriminator_forward
        try:
            yield
        finally:
            _discriminator.Discriminator.forward = old


# ----------------------------------------------------------------------------------------------------------------------
# The MIRROR's whole COMPRESSION_GAN training iteration on the CPU (tests/test_train_ddp.py): on top of the drop-in
# stand-ins, the mirror's own loss modules -- LPIPS (fused per-layer kernels + cuDNN trunk on the GPU) and the fused GAN
# loss -- are replaced by the oracle's torch formulas.
# ----------------------------------------------------------------------------------------------------------------------
from hific_b200.loss import perceptual as _perceptual


def perceptual_forward(self, pred, target, normalize=False):
    return O.lpips_forward(self.trunk, [w for w in self.lins], pred, target, normalize=normalize)


def gan_loss_fn(logits_real, logits_gen, mode):
    d_loss, g_loss = O.gan_losses_non_saturating(logits_real, logits_gen)
    return d_loss if mode == 1 else g_loss


@contextlib.contextmanager
def gan_model_cpu_emulation():
    saved = []

    def patch(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    with dropin_cpu_emulation():
        patch(_perceptual.PerceptualLoss, "forward", perceptual_forward)
        patch(ops.GanLossFn, "apply", staticmethod(gan_loss_fn))
        try:
            yield
        finally:
            for obj, name, value in reversed(saved):
                setattr(obj, name, value)
