"""Training-mode (autograd) execution of the HiFIC networks on the B200 kernels.

The inference plans in `engine.py` fuse ChannelNorm into conv epilogues and recycle buffers.  Training needs the
pre-norm conv outputs and every layer input for the backward pass, so the training plans run each layer un-fused
(conv -> fp32 rows -> ChannelNorm kernel -> fp16 act buffer), keep all of it, and walk the same list backwards:
ChannelNorm backward -> weight / bias gradient (tcgen05 GEMM over the pixels) -> data gradient (tcgen05 conv on the
gradient).  Each network is exposed to torch as ONE autograd.Function (inputs: boundary tensor + parameters), so
`loss.backward()` and the reference's optimizers (train.py:49-59, 287-300) work unchanged.

Reference autograd paths replaced: Encoder (src/network/encoder.py:104-111), Generator (generator.py:145-169),
HyperpriorAnalysis / HyperpriorSynthesis (hyper.py:56-63, 90-97).
"""
import ctypes

import torch

from . import ops
from ._lib import check, lib
from . import grad as _grad
from .grad import GRAD_BF16, ConvGrad, GradScale, Workspace
from .ops import (ACT_NONE, ACT_RELU, OUT_NCHW_F32, OUT_NHWC_F16, OUT_NHWC_F32, PAD_REFLECT, PAD_ZERO, CN_EPS, Conv,
                  Geom, _ptr, _stream, round_up)


def nchw_to_rows(t):
    """(n, c, h, w) fp32 -> fp32 rows [n*h*w][round_up(c, 4)] (pure data movement)."""
    n, c, h, w = t.shape
    r = t.permute(0, 2, 3, 1).reshape(-1, c)
    if c % 4:
        r = torch.nn.functional.pad(r, (0, 4 - c % 4))
    return r.contiguous()


def rows_to_nchw(r, n, c, h, w):
    return r.view(n, h, w, -1)[..., :c].permute(0, 3, 1, 2).contiguous()


def norm_bwd(z, g, gamma, beta, act, as_operand=True, kind="channel", n=None):
    """ChannelNorm / InstanceNorm (+ReLU) backward on fp32 rows (kind "instance": n = images, the rows are n * hw).  Returns (dz, dgamma, dbeta, dbias) with dbias = column sums of dz
    (the gradient of the bias of the convolution that produced z).  as_operand: dz is written directly as the bf16
    NHWC operand of the two backward GEMMs of that convolution (pitch round_up(c, 64), shared workspace -- consume it
    before the next norm_bwd); otherwise as fp32 rows."""
    c = gamma.numel()
    npix = z.shape[0]
    dgb = torch.zeros((3, c), dtype=torch.float32, device=z.device)
    if as_operand:
        cpad = round_up(c, 64)
        dz = Workspace.get("dy_act", npix * cpad, torch.int16, z.device).view(npix, cpad)
        dz_f32, dz_act, ld = None, dz, 0
    else:
        dz = torch.empty((npix, round_up(c, 4)), dtype=torch.float32, device=z.device)
        dz_f32, dz_act, ld, cpad = dz, None, dz.shape[1], 0
    if kind == "instance":
        ws, ws_bytes = ops.instancenorm_ws(n, c, z.device)
        check(lib.hfc_instancenorm_bwd(_ptr(z), z.shape[1], _ptr(g), g.shape[1], _ptr(gamma.detach().reshape(-1)),
                                       _ptr(beta.detach().reshape(-1)), c, n, npix // n, ops.IN_EPS, act, _ptr(dz_f32), ld,
                                       _ptr(dgb[0]), _ptr(dgb[1]), _ptr(dgb[2]), _ptr(dz_act), cpad, int(GRAD_BF16),
                                       _ptr(ws), ws_bytes, _stream()), "instancenorm_bwd")
    else:
        check(lib.hfc_channelnorm_bwd(_ptr(z), z.shape[1], _ptr(g), g.shape[1], _ptr(gamma.detach().reshape(-1)),
                                      _ptr(beta.detach().reshape(-1)), c, npix, CN_EPS, act, _ptr(dz_f32), ld,
                                      _ptr(dgb[0]), _ptr(dgb[1]), _ptr(dgb[2]), _ptr(dz_act), cpad, int(GRAD_BF16), _stream()),
              "channelnorm_bwd")
    if as_operand:
        _grad._log_operand(dz)
    if _grad._inv_scale != 1.0:
        dgb.mul_(_grad._inv_scale)           # parameter gradients leave the layer un-scaled (grad.GradScale); dz keeps S
    return dz, dgb[0].view_as(gamma), dgb[1].view_as(beta), dgb[2]


def relu_mask(g, y_act, geom, slope=0.0):
    """g * act'(.) from the layer's saved fp16 output: ReLU (slope 0) or LeakyReLU(slope)."""
    out = torch.empty((g.shape[0], round_up(geom.c, 4)), dtype=torch.float32, device=g.device)
    if out.shape[1] != geom.c:
        out.zero_()
    gs = geom.c_struct()
    check(lib.hfc_relu_mask(_ptr(g), g.shape[1], _ptr(y_act), ctypes.byref(gs), float(slope), _ptr(out), out.shape[1],
                             _stream()),
          "relu_mask")
    return out


class Layer:
    """One conv / transposed conv of a training plan with its forward kernel, its backward and its saved tensors."""

    def __init__(self, in_geom, cout, k, stride=1, transposed=False, pad_mode=PAD_ZERO, pad=(0, 0, 0, 0), window=False,
                 fused_relu_geom=None, nchw_out=False, fused_act=ACT_RELU):
        self.in_geom, self.cout = in_geom, cout
        kw = dict(stride=stride, transposed=transposed, pad_mode=pad_mode, pad=pad, window=window)
        if nchw_out:
            self.conv = Conv(in_geom, cout, k, out_mode=OUT_NCHW_F32, **kw)
        elif fused_relu_geom is not None:    # bias + ReLU in the epilogue, bordered fp16 output (hyper networks)
            self.conv = Conv(in_geom, cout, k, out_geom=fused_relu_geom, out_reflect=any(
                (fused_relu_geom.pt, fused_relu_geom.pl, fused_relu_geom.pb, fused_relu_geom.pr)), act=fused_act, **kw)
        else:
            og = self.conv_out_geom(in_geom, cout, k, stride, transposed, pad)
            self.conv = Conv(in_geom, cout, k, out_mode=OUT_NHWC_F32, out_geom=og, **kw)
        self.grad = ConvGrad(in_geom, cout, k, stride=stride, transposed=transposed, pad_mode=pad_mode, pad=pad)
        self.oh, self.ow = self.grad.oh, self.grad.ow
        self.x_act = None

    @staticmethod
    def conv_out_geom(in_geom, cout, k, stride, transposed, pad):
        n, h, w = in_geom.n, in_geom.h, in_geom.w
        if transposed:
            oh, ow = (h - 1) * stride - 2 * pad[0] + k + stride - 1, (w - 1) * stride - 2 * pad[1] + k + stride - 1
        else:
            oh, ow = (h + pad[0] + pad[2] - k) // stride + 1, (w + pad[1] + pad[3] - k) // stride + 1
        return Geom(n, oh, ow, cout, round_up(cout, 4))

    def forward(self, x_act, weight, bias):
        self.x_act = x_act
        return self.conv(x_act, weight, bias)

    def backward(self, dz, weight, need_dx=True, db=None):
        """dz: fp32 gradient rows, or (int16 = bf16 bits) the operand buffer written by norm_bwd(as_operand=True)."""
        if dz.dtype == torch.int16:
            dy_act, dz_rows = dz.view(self.grad.dy_geom.shape), None
            assert db is not None
        else:
            dy_act, dz_rows = self.grad.dy_to_act(dz), dz       # one bf16 copy of the gradient serves both GEMMs
        dw = self.grad.weight_grad(self.x_act, dz_rows, dy_act=dy_act)
        if db is None:
            db = self.grad.bias_grad(dz_rows)
        dx = self.grad.data_grad(dz_rows, weight.detach(), dy_act=dy_act) if need_dx else None
        return dx, dw, db


def norm_fwd(z, geom, gamma, beta, act, reflect, res1=None, res2=None, want_f32=False, want_act=True, kind="channel"):
    """The inter-layer normalisation of the plan: ChannelNorm2D (default) or InstanceNorm2d (use_channel_norm=False)."""
    fn = ops.instancenorm if kind == "instance" else ops.channelnorm
    return fn(z, geom, gamma, beta, act=act, reflect=reflect, res1=res1, res2=res2, want_f32=want_f32, want_act=want_act)


# ------------------------------------------------------------------------------------------------------------
# Encoder
# ------------------------------------------------------------------------------------------------------------
class EncoderTrainPlan:
    FILTERS = (60, 120, 240, 480, 960)

    def __init__(self, n, h, w, im_channels, C, device, norm_kind="channel"):
        f = self.FILTERS
        self.n = n
        self.kind = norm_kind
        self.g_in = Geom(n, h, w, im_channels, 8, 3, 3, 3, 4)
        asym = (1, 0, 0, 1)
        self.layers, self.geoms = [], []
        self.layers.append(Layer(self.g_in, f[0], 7, pad_mode=PAD_REFLECT, pad=(3, 3, 3, 3), window=True))
        g = Geom(n, h, w, f[0], 64, *asym)
        self.geoms.append(g)
        for i in range(1, 5):
            self.layers.append(Layer(g, f[i], 3, stride=2, pad_mode=PAD_REFLECT, pad=(1, 0, 0, 1)))
            lay = self.layers[-1]
            g = Geom(n, lay.oh, lay.ow, f[i], round_up(f[i], 64), *(asym if i < 4 else (1, 1, 1, 1)))
            self.geoms.append(g)
        self.out = Layer(g, C, 3, pad_mode=PAD_REFLECT, pad=(1, 1, 1, 1), nchw_out=True)
        self.C = C

    def forward(self, x, p):
        """p: flat parameter list in module.parameters() order: (w, b, gamma, beta) x 5, (w_out, b_out)."""
        self.z = []
        h = ops.nchw_to_act(x, self.g_in, reflect=True)
        for i, lay in enumerate(self.layers):
            w, b, gm, bt = p[4 * i:4 * i + 4]
            z = lay.forward(h, w, b)
            self.z.append(z)
            h, _ = norm_fwd(z, self.geoms[i], gm, bt, ACT_RELU, True, kind=self.kind)
        return self.out.forward(h, p[20], p[21])

    def backward(self, dy, p):
        grads = [None] * 22
        g, grads[20], grads[21] = self.out.backward(nchw_to_rows(dy), p[20])
        _grad.emit(grads[20], grads[21])
        for i in range(4, -1, -1):
            w, b, gm, bt = p[4 * i:4 * i + 4]
            dz, grads[4 * i + 2], grads[4 * i + 3], db = norm_bwd(self.z[i], g, gm, bt, ACT_RELU, kind=self.kind, n=self.n)
            g, grads[4 * i], grads[4 * i + 1] = self.layers[i].backward(dz, w, need_dx=i > 0, db=db)
            _grad.emit(*grads[4 * i:4 * i + 4])
        return grads

    def release(self):
        self.z = None


# ------------------------------------------------------------------------------------------------------------
# Generator
# ------------------------------------------------------------------------------------------------------------
class GeneratorTrainPlan:
    FILTERS = (960, 480, 240, 120, 60)

    def __init__(self, n, h, w, C, n_res, im_channels, device, noise_dim=0, norm_kind="channel"):
        f = self.FILTERS
        self.n, self.h, self.w, self.C, self.n_res = n, h, w, C, n_res
        self.kind = norm_kind
        # sample_noise (generator.py:105-107, 149-153): the trunk is F0 = 960 + noise_dim channels wide
        self.noise_dim, self.F0 = noise_dim, f[0] + noise_dim
        F0, F0p = self.F0, round_up(self.F0, 64)
        b1 = (1, 1, 1, 1)
        self.g_in = Geom(n, h, w, C, round_up(C, 64), *b1)
        self.g_head = Geom(n, h, w, 960, 960)
        self.g_b1 = Geom(n, h, w, F0, F0p, *b1)
        self.g_flat = Geom(n, h, w, F0, F0p)
        self.init = Layer(self.g_in, 960, 3, pad_mode=PAD_REFLECT, pad=b1)
        self.res = [(Layer(self.g_b1, F0, 3, pad_mode=PAD_REFLECT, pad=b1), Layer(self.g_b1, F0, 3, pad_mode=PAD_REFLECT, pad=b1))
                    for _ in range(n_res)]
        self.ups, self.up_geoms = [], []
        g = self.g_flat
        for i in range(1, 5):
            lay = Layer(g, f[i], 3, stride=2, transposed=True, pad=(1, 1, 1, 1))
            g = Geom(n, lay.oh, lay.ow, f[i], round_up(f[i], 64), *((0, 0, 0, 0) if i < 4 else (3, 3, 3, 3)))
            self.ups.append(lay)
            self.up_geoms.append(g)
        self.out = Layer(g, im_channels, 7, pad_mode=PAD_REFLECT, pad=(3, 3, 3, 3), nchw_out=True)

    # parameter order of hific_b200.network.generator.Generator.parameters():
    #   init: norm0 (gamma, beta), conv (w, b), norm3 (gamma, beta)                      -> 0..5
    #   resblock m: conv1 (w, b), conv2 (w, b), norm1 (gamma, beta), norm2 (gamma, beta) -> 6 + 8m ..
    #   upconv i: convT (w, b), norm (gamma, beta)                                       -> 6 + 8R + 4i ..
    #   out: conv (w, b)
    def forward(self, y_hat, p):
        R = self.n_res
        self.y_rows = nchw_to_rows(y_hat)
        k = dict(kind=self.kind)
        a0, _ = norm_fwd(self.y_rows, self.g_in, p[0], p[1], ACT_NONE, True, **k)
        self.z_init = self.init.forward(a0, p[2], p[3])
        if self.noise_dim == 0:
            x_act, head = norm_fwd(self.z_init, self.g_b1 if R else self.g_flat, p[4], p[5], ACT_NONE, bool(R), want_f32=True, **k)
        else:
            # head = cat(norm(conv(.)), z), z ~ N(0, 1) (generator.py:150-153): data movement with torch ops; the noise
            # channels carry no gradient back
            _, head960 = norm_fwd(self.z_init, self.g_head, p[4], p[5], ACT_NONE, False, want_f32=True, want_act=False, **k)
            z = torch.randn((self.n, self.noise_dim, self.h, self.w)).to(head960)
            hn = torch.cat((head960.view(self.n, self.h, self.w, 960).permute(0, 3, 1, 2), z), dim=1).contiguous()
            head = hn.permute(0, 2, 3, 1).reshape(-1, self.F0).contiguous()
            x_act = ops.nchw_to_act(hn, self.g_b1 if R else self.g_flat, reflect=bool(R))
        self.zr = []
        x_f32 = head
        for m in range(R):
            w1, bb1, w2, bb2, g1, be1, g2, be2 = p[6 + 8 * m:14 + 8 * m]
            c1, c2 = self.res[m]
            z1 = c1.forward(x_act, w1, bb1)
            a1, _ = norm_fwd(z1, self.g_b1, g1, be1, ACT_RELU, True, **k)
            z2 = c2.forward(a1, w2, bb2)
            last = m == R - 1
            x_act, x_new = norm_fwd(z2, self.g_flat if last else self.g_b1, g2, be2, ACT_NONE, not last, res1=x_f32,
                                    res2=head if last else None, want_f32=not last, **k)
            self.zr.append((z1, z2))
            x_f32 = x_new
        self.zu = []
        h = x_act
        o = 6 + 8 * R
        for i, lay in enumerate(self.ups):
            w, b, gm, bt = p[o + 4 * i:o + 4 * i + 4]
            z = lay.forward(h, w, b)
            self.zu.append(z)
            gg = self.up_geoms[i]
            h, _ = norm_fwd(z, gg, gm, bt, ACT_RELU, any((gg.pt, gg.pl, gg.pb, gg.pr)), **k)
        return self.out.forward(h, p[o + 16], p[o + 17])

    def backward(self, dxhat, p):
        R = self.n_res
        o = 6 + 8 * R
        grads = [None] * len(p)
        k = dict(kind=self.kind, n=self.n)
        g, grads[o + 16], grads[o + 17] = self.out.backward(nchw_to_rows(dxhat), p[o + 16])
        _grad.emit(grads[o + 16], grads[o + 17])
        for i in range(3, -1, -1):
            w, b, gm, bt = p[o + 4 * i:o + 4 * i + 4]
            dz, grads[o + 4 * i + 2], grads[o + 4 * i + 3], db = norm_bwd(self.zu[i], g, gm, bt, ACT_RELU, **k)
            g, grads[o + 4 * i], grads[o + 4 * i + 1] = self.ups[i].backward(dz, w, db=db)
            _grad.emit(*grads[o + 4 * i:o + 4 * i + 4])
        # g = gradient w.r.t. the trunk output x_R (+ head): identity paths carry it to every block input and to head
        g_head = g.clone() if R else g
        for m in range(R - 1, -1, -1):
            w1, bb1, w2, bb2, g1, be1, g2, be2 = p[6 + 8 * m:14 + 8 * m]
            c1, c2 = self.res[m]
            z1, z2 = self.zr[m]
            dz2, grads[6 + 8 * m + 6], grads[6 + 8 * m + 7], db2 = norm_bwd(z2, g, g2, be2, ACT_NONE, **k)
            ga1, grads[6 + 8 * m + 2], grads[6 + 8 * m + 3] = c2.backward(dz2, w2, db=db2)
            _grad.emit(grads[6 + 8 * m + 2], grads[6 + 8 * m + 3], grads[6 + 8 * m + 6], grads[6 + 8 * m + 7])
            dz1, grads[6 + 8 * m + 4], grads[6 + 8 * m + 5], db1 = norm_bwd(z1, ga1, g1, be1, ACT_RELU, **k)
            gx, grads[6 + 8 * m], grads[6 + 8 * m + 1] = c1.backward(dz1, w1, db=db1)
            _grad.emit(grads[6 + 8 * m], grads[6 + 8 * m + 1], grads[6 + 8 * m + 4], grads[6 + 8 * m + 5])
            g = g + gx                                  # identity_map + residual branch (generator.py:44)
        if R:
            g_head = g_head + g                         # block 0 consumed head; the final `x += head` added it again
        if self.noise_dim:
            g_head = g_head[:, :960].contiguous()       # the noise channels of the head are constants
        dz0, grads[4], grads[5], db0 = norm_bwd(self.z_init, g_head, p[4], p[5], ACT_NONE, **k)
        ga0, grads[2], grads[3] = self.init.backward(dz0, p[2], db=db0)
        dy_rows, grads[0], grads[1], _ = norm_bwd(self.y_rows, ga0, p[0], p[1], ACT_NONE, as_operand=False, **k)
        _grad.emit(*grads[0:6])
        return rows_to_nchw(dy_rows, self.n, self.C, self.h, self.w), grads

    def release(self):
        self.zr = self.zu = self.z_init = self.y_rows = None


class ResidualBlockTrainPlan:
    """One ResidualBlock.forward (src/network/generator.py:33-44) called on its own: pad, conv, norm, ReLU, pad, conv,
    norm, + identity.  Serves both the no-grad call and autograd (PlanFunction); inside Generator.forward the blocks run
    as part of the Generator plans instead.  Parameter order: conv1 (w, b), conv2 (w, b), norm1 (g, b), norm2 (g, b)."""

    def __init__(self, n, h, w, c, device, norm_kind="channel"):
        self.n, self.h, self.w, self.c = n, h, w, c
        self.kind = norm_kind
        b1 = (1, 1, 1, 1)
        self.g_b1 = Geom(n, h, w, c, round_up(c, 64), *b1)
        self.g_flat = Geom(n, h, w, c, round_up(c, 64))
        self.c1 = Layer(self.g_b1, c, 3, pad_mode=PAD_REFLECT, pad=b1)
        self.c2 = Layer(self.g_b1, c, 3, pad_mode=PAD_REFLECT, pad=b1)

    def forward(self, x, p):
        w1, bb1, w2, bb2, g1, be1, g2, be2 = p
        x_rows = nchw_to_rows(x)
        x_act = ops.nchw_to_act(x, self.g_b1, reflect=True)
        self.z1 = self.c1.forward(x_act, w1, bb1)
        a1, _ = norm_fwd(self.z1, self.g_b1, g1, be1, ACT_RELU, True, kind=self.kind)
        self.z2 = self.c2.forward(a1, w2, bb2)
        _, out = norm_fwd(self.z2, self.g_flat, g2, be2, ACT_NONE, False, res1=x_rows, want_f32=True, want_act=False,
                          kind=self.kind)
        return rows_to_nchw(out, self.n, self.c, self.h, self.w)

    def backward(self, dout, p):
        w1, bb1, w2, bb2, g1, be1, g2, be2 = p
        grads = [None] * 8
        g = nchw_to_rows(dout)
        k = dict(kind=self.kind, n=self.n)
        dz2, grads[6], grads[7], db2 = norm_bwd(self.z2, g, g2, be2, ACT_NONE, **k)
        ga1, grads[2], grads[3] = self.c2.backward(dz2, w2, db=db2)
        dz1, grads[4], grads[5], db1 = norm_bwd(self.z1, ga1, g1, be1, ACT_RELU, **k)
        gx, grads[0], grads[1] = self.c1.backward(dz1, w1, db=db1)
        _grad.emit(*grads)
        return rows_to_nchw(g[:, :self.c] + gx[:, :self.c], self.n, self.c, self.h, self.w), grads

    def release(self):
        self.z1 = self.z2 = None


# ------------------------------------------------------------------------------------------------------------
# Hyper networks (bias + ReLU fused in the conv epilogue; the saved fp16 outputs give the ReLU masks)
# ------------------------------------------------------------------------------------------------------------
class HyperAnalysisTrainPlan:
    def __init__(self, n, h, w, C, N, device):
        self.n, self.h, self.w, self.C, self.N = n, h, w, C, N
        b2 = (2, 2, 2, 2)
        self.g_in = Geom(n, h, w, C, round_up(C, 64))
        self.g1 = Geom(n, h, w, N, round_up(N, 64), *b2)
        self.l1 = Layer(self.g_in, N, 3, pad_mode=PAD_ZERO, pad=(1, 1, 1, 1), fused_relu_geom=self.g1)
        h2, w2 = (h + 4 - 5) // 2 + 1, (w + 4 - 5) // 2 + 1
        self.g2 = Geom(n, h2, w2, N, round_up(N, 64), *b2)
        self.l2 = Layer(self.g1, N, 5, stride=2, pad_mode=PAD_REFLECT, pad=b2, fused_relu_geom=self.g2)
        self.l3 = Layer(self.g2, N, 5, stride=2, pad_mode=PAD_REFLECT, pad=b2, nchw_out=True)

    def forward(self, y, p):
        a0 = ops.nchw_to_act(y, self.g_in)
        self.a1 = self.l1.forward(a0, p[0], p[1])
        self.a2 = self.l2.forward(self.a1, p[2], p[3])
        return self.l3.forward(self.a2, p[4], p[5])

    def backward(self, dz, p):
        grads = [None] * 6
        g, grads[4], grads[5] = self.l3.backward(nchw_to_rows(dz), p[4])
        _grad.emit(grads[4], grads[5])
        g = relu_mask(g, self.a2, self.g2)
        g, grads[2], grads[3] = self.l2.backward(g, p[2])
        _grad.emit(grads[2], grads[3])
        g = relu_mask(g, self.a1, self.g1)
        g, grads[0], grads[1] = self.l1.backward(g, p[0])
        _grad.emit(grads[0], grads[1])
        return rows_to_nchw(g, self.n, self.C, self.h, self.w), grads

    def release(self):
        self.a1 = self.a2 = None


class HyperSynthesisTrainPlan:
    def __init__(self, n, h, w, C, N, device):
        self.n, self.h, self.w, self.C, self.N = n, h, w, C, N
        self.g_in = Geom(n, h, w, N, round_up(N, 64))
        self.g1 = Geom(n, 2 * h, 2 * w, N, round_up(N, 64))
        self.g2 = Geom(n, 4 * h, 4 * w, N, round_up(N, 64))
        self.l1 = Layer(self.g_in, N, 5, stride=2, transposed=True, pad=(2, 2, 2, 2), fused_relu_geom=self.g1)
        self.l2 = Layer(self.g1, N, 5, stride=2, transposed=True, pad=(2, 2, 2, 2), fused_relu_geom=self.g2)
        self.l3 = Layer(self.g2, C, 3, stride=1, transposed=True, pad=(1, 1, 1, 1), nchw_out=True)

    def forward(self, z, p):
        a0 = ops.nchw_to_act(z, self.g_in)
        self.a1 = self.l1.forward(a0, p[0], p[1])
        self.a2 = self.l2.forward(self.a1, p[2], p[3])
        return self.l3.forward(self.a2, p[4], p[5])

    def backward(self, dout, p):
        grads = [None] * 6
        g, grads[4], grads[5] = self.l3.backward(nchw_to_rows(dout), p[4])
        _grad.emit(grads[4], grads[5])
        g = relu_mask(g, self.a2, self.g2)
        g, grads[2], grads[3] = self.l2.backward(g, p[2])
        _grad.emit(grads[2], grads[3])
        g = relu_mask(g, self.a1, self.g1)
        g, grads[0], grads[1] = self.l1.backward(g, p[0])
        _grad.emit(grads[0], grads[1])
        return rows_to_nchw(g, self.n, self.N, self.h, self.w), grads

    def release(self):
        self.a1 = self.a2 = None


class HyperSynthesisDLMMTrainPlan:
    """HyperpriorSynthesisDLMM (src/network/hyper.py:100-130): two ReLU transposed convs, a LINEAR transposed conv kept as
    an fp16 activation buffer, and the linear 1x1 conv to the 3*K*C mixture parameters (NCHW fp32)."""

    def __init__(self, n, h, w, C, N, n_out, device):
        self.n, self.h, self.w, self.C, self.N = n, h, w, C, N
        self.g_in = Geom(n, h, w, N, round_up(N, 64))
        self.g1 = Geom(n, 2 * h, 2 * w, N, round_up(N, 64))
        self.g2 = Geom(n, 4 * h, 4 * w, N, round_up(N, 64))
        self.g3 = Geom(n, 4 * h, 4 * w, C, round_up(C, 64))
        self.l1 = Layer(self.g_in, N, 5, stride=2, transposed=True, pad=(2, 2, 2, 2), fused_relu_geom=self.g1)
        self.l2 = Layer(self.g1, N, 5, stride=2, transposed=True, pad=(2, 2, 2, 2), fused_relu_geom=self.g2)
        self.l3 = Layer(self.g2, C, 3, stride=1, transposed=True, pad=(1, 1, 1, 1), fused_relu_geom=self.g3,
                        fused_act=ACT_NONE)
        self.l4 = Layer(self.g3, n_out, 1, nchw_out=True)

    def forward(self, z, p):
        a0 = ops.nchw_to_act(z, self.g_in)
        self.a1 = self.l1.forward(a0, p[0], p[1])
        self.a2 = self.l2.forward(self.a1, p[2], p[3])
        a3 = self.l3.forward(self.a2, p[4], p[5])
        return self.l4.forward(a3, p[6], p[7])

    def backward(self, dout, p):
        grads = [None] * 8
        g, grads[6], grads[7] = self.l4.backward(nchw_to_rows(dout), p[6])
        g, grads[4], grads[5] = self.l3.backward(g, p[4])                      # linear layer: no mask
        _grad.emit(grads[4], grads[5], grads[6], grads[7])
        g = relu_mask(g, self.a2, self.g2)
        g, grads[2], grads[3] = self.l2.backward(g, p[2])
        _grad.emit(grads[2], grads[3])
        g = relu_mask(g, self.a1, self.g1)
        g, grads[0], grads[1] = self.l1.backward(g, p[0])
        _grad.emit(grads[0], grads[1])
        return rows_to_nchw(g, self.n, self.N, self.h, self.w), grads

    def release(self):
        self.a1 = self.a2 = None


# ------------------------------------------------------------------------------------------------------------
# Discriminator: the forward IS the inference plan (engine.DiscriminatorPlan keeps every layer output in its own
# buffers: bias + LeakyReLU fused, 1/sigma folded into the weight packing); the backward walks it in reverse.
# ------------------------------------------------------------------------------------------------------------
class DiscriminatorTrainPlan:
    """Autograd of Discriminator.forward (src/network/discriminator.py:66-86) incl. the spectral-norm
    reparametrisation W = W_orig / sigma (u, v constants, as torch.nn.utils.spectral_norm).
    Parameter order: context_conv (w, b), conv1..4 (weight_orig, bias), conv_out (w, b).
    The context latents are detached by the caller (src/model.py:178), so no gradient flows into `y`."""
    SLOPE = 0.2

    def __init__(self, fwd_plan):
        from .engine import DiscriminatorPlan
        assert isinstance(fwd_plan, DiscriminatorPlan)
        self.f = fwd_plan
        b1 = (1, 1, 1, 1)
        self.g_ctx = ConvGrad(fwd_plan.g_y, fwd_plan.CONTEXT_C, 3, pad_mode=PAD_REFLECT, pad=b1)
        self.grads = [ConvGrad(c.in_geom, f, 4, stride=2, pad_mode=PAD_REFLECT, pad=b1)
                      for c, f in zip(fwd_plan.convs, fwd_plan.FILTERS)]
        self.g_out = ConvGrad(fwd_plan.c_out.in_geom, 1, 1)
        self.out_geoms = [c.out_geom for c in fwd_plan.convs]

    def forward(self, mod, x, y):
        f = self.f
        logits = f.run(mod, x, y)
        # constants of the spectral-norm backward as they were in THIS forward (u, v are updated in place by the next)
        self.sn = [(getattr(mod, f"conv{i + 1}").weight_u.clone(), getattr(mod, f"conv{i + 1}").weight_v.clone(),
                    f.inv_sigmas[i]) for i in range(4)]
        return logits

    def backward(self, dlogits, p, need_dx):
        f = self.f
        dev = dlogits.device
        grads = [None] * 12
        g, grads[10], grads[11] = self._layer_bwd(self.g_out, f.bufs[3], nchw_to_rows(dlogits), p[10], None)
        for i in range(3, -1, -1):
            g = relu_mask(g, f.bufs[i], self.out_geoms[i], self.SLOPE)
            x_in = f.bufs[i - 1] if i else f.in_act
            u, v, inv_sigma = self.sn[i]
            w_orig = p[2 + 2 * i]
            g_next, dw, grads[3 + 2 * i] = self._layer_bwd(self.grads[i], x_in, g, w_orig, inv_sigma, need_dx=True)
            dwo = torch.empty_like(w_orig)
            ws = torch.empty(1, dtype=torch.float32, device=dev)
            rows, cols = w_orig.shape[0], w_orig.numel() // w_orig.shape[0]
            check(lib.hfc_spectral_bwd(_ptr(dw), _ptr(w_orig), _ptr(u), _ptr(v), _ptr(inv_sigma), rows, cols, _ptr(ws), 0,
                                       _ptr(dwo), _stream()), "spectral_bwd")
            grads[2 + 2 * i] = dwo
            g = g_next
        gi = f.g_in
        dx = torch.empty((gi.n, f.c_x, gi.h, gi.w), dtype=torch.float32, device=dev) if need_dx else None
        gc = f.g_ctx
        dctx = torch.empty((gc.n * gc.h * gc.w, round_up(gc.c, 4)), dtype=torch.float32, device=dev)
        check(lib.hfc_disc_input_bwd(_ptr(g), g.shape[1], gi.n, gi.h, gi.w, f.c_x, gc.c, f.scale, _ptr(dx), _ptr(dctx),
                                     dctx.shape[1], _stream()), "disc_input_bwd")
        dctx = relu_mask(dctx, f.ctx_act, gc, self.SLOPE)
        grads[0] = self.g_ctx.weight_grad(f.y_act, dctx)
        grads[1] = self.g_ctx.bias_grad(dctx)
        return dx, grads

    def release(self):
        self.sn = None

    @staticmethod
    def _layer_bwd(cg, x_act, dz_rows, weight, inv_sigma, need_dx=True):
        dy_act = cg.dy_to_act(dz_rows)
        dw = cg.weight_grad(x_act, dz_rows, dy_act=dy_act)
        db = cg.bias_grad(dz_rows)
        dx = cg.data_grad(dz_rows, weight.detach(), scale=inv_sigma, dy_act=dy_act) if need_dx else None
        return dx, dw, db


def _stamp(plan):
    """Saved activations live on the (cached, shared) plan, not on the autograd ctx: every forward stamps the plan with a
    new generation and its backward refuses to run against a plan that another forward has overwritten since (two
    forwards of the same shape before one backward -- micro-batch accumulation with a summed loss -- or a no-grad
    Discriminator forward in between, which shares the plan's buffers)."""
    plan._generation = getattr(plan, "_generation", 0) + 1
    return plan._generation


def _check_stamp(plan, generation, who):
    if getattr(plan, "_generation", None) != generation:
        raise RuntimeError(f"{who}: the plan's saved activations were overwritten by a later forward of the same shape "
                           "before this backward ran (hific_b200 keeps ONE set of saved tensors per input shape: call "
                           "backward() before the next forward of that shape, or accumulate gradients step by step)")


class DiscriminatorFunction(torch.autograd.Function):
    """forward(plan, module, x, y, *params) -> logits; y gets no gradient (the reference detaches the latents)."""

    @staticmethod
    def forward(ctx, plan, mod, x, y, *params):
        ctx.plan, ctx.params, ctx.needs_dx = plan, params, x.requires_grad
        ctx.generation = _stamp(plan)
        with torch.no_grad():
            out = plan.forward(mod, x.contiguous(), y.contiguous())
        ctx.generation_f = plan.f._generation                      # the inference plan owns the shared buffers
        return out

    @staticmethod
    def backward(ctx, dlogits):
        plan = ctx.plan
        _check_stamp(plan, ctx.generation, "Discriminator backward")
        _check_stamp(plan.f, ctx.generation_f, "Discriminator backward")
        if not hasattr(plan, "grad_scale"):
            plan.grad_scale = GradScale()
        with torch.no_grad():
            params = [q.detach() for q in ctx.params]
            dx, grads = plan.grad_scale.run(lambda d: plan.backward(d.contiguous(), params, ctx.needs_dx), dlogits)
            plan.release()
        grads = [g.reshape(q.shape) if g is not None else None for g, q in zip(grads, ctx.params)]
        return (None, None, dx, None, *grads)


# ------------------------------------------------------------------------------------------------------------
# autograd glue: one Function per network
# ------------------------------------------------------------------------------------------------------------
class PlanFunction(torch.autograd.Function):
    """forward(plan, x, *params) -> plan.forward(x, params); backward walks the plan in reverse."""

    @staticmethod
    def forward(ctx, plan, x, *params):
        ctx.plan = plan
        ctx.params = params
        ctx.needs_dx = x.requires_grad
        ctx.generation = _stamp(plan)
        with torch.no_grad():
            return plan.forward(x.contiguous(), [q.detach() for q in params])

    @staticmethod
    def backward(ctx, dout):
        plan = ctx.plan
        _check_stamp(plan, ctx.generation, type(plan).__name__ + " backward")
        if not hasattr(plan, "grad_scale"):
            plan.grad_scale = GradScale()

        def walk(d):
            res = plan.backward(d.contiguous(), params)
            return res if isinstance(res, tuple) else (None, res)
        with torch.no_grad():
            params = [q.detach() for q in ctx.params]
            dx, grads = plan.grad_scale.run(walk, dout)
            plan.release()
            if _grad._grad_sink is not None:
                _grad._grad_sink.function_done()           # autograd reads the returned tensors next: reductions must be in
        grads = [g.reshape(q.shape) if g is not None else None for g, q in zip(grads, ctx.params)]
        return (None, dx if ctx.needs_dx else None, *grads)


def run_training(plan, x, params):
    return PlanFunction.apply(plan, x, *params)


def run_inference(plan, x, params):
    """The plan's forward without autograd (variants that have no fused inference plan: InstanceNorm, a stand-alone
    ResidualBlock): the layer-by-layer forward, saved tensors dropped at once."""
    _stamp(plan)                      # a backward still pending on this plan must fail loudly, not read these buffers
    with torch.no_grad():
        out = plan.forward(x.contiguous(), [q.detach() for q in params])
        plan.release()
    return out
