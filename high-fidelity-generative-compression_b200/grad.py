"""Backward of one convolution / transposed convolution on the B200 kernels (autograd of F.conv2d /
F.conv_transpose2d as issued by the reference's training step, train.py:49-59).

  data gradient    the same tcgen05 implicit-GEMM kernel run on the gradient tensor:
                     conv stride 1          -> stride-1 conv with transposed + flipped weights over the PADDED input
                                               domain, then the adjoint of the (reflect / zero) padding (hfc_pad_fold)
                     conv stride 2          -> transposed conv (4 sub-pixel phases) with the forward weights
                     transposed conv        -> (strided) conv with the forward weights
  weight gradient  one tcgen05 GEMM over K = pixels: dW[c1][(tap, c2)] = A1T . COLT^T with both operands built by
                   the tiled transposes in csrc/backward.cu (explicit transposed im2col: simple and correct; an
                   implicit wgrad kernel is the obvious next optimisation)
  bias gradient    column sums of the gradient rows

Gradients travel as fp32 rows [pixels][channels].  Operand format of the backward GEMMs (process-wide, HFC_GRAD_FMT):

  fp16 (default)  the gradient operand is rounded to fp16 -- the 10-bit mantissa of the TF32 path cuDNN gives the reference
                  on a GPU -- and meets the fp16 activations saved by the forward pass and the fp16 weights as they are
                  (no conversion pass).  fp16's narrow exponent is handled by a power-of-two LOSS SCALE per backward
                  Function (`GradScale`): the incoming gradient is multiplied by S, everything in between is linear, the
                  parameter / input gradients are divided by S on the way out (exact).  Conversions saturate at +-65504
                  instead of producing inf, so a stale scale clips outliers rather than poisoning the step.
  bf16            round 1's path: bf16 gradients (fp32 exponent range, no scaling, 8-bit mantissa) against activations and
                  weights converted to bf16 on the fly; kept for the side-by-side precision study (tools/grad_precision.py).
"""
import ctypes
import os

import torch

from . import _lib, ops
from ._lib import check, lib
from .ops import OUT_NHWC_F32, PAD_REFLECT, PAD_ZERO, Conv, Geom, _ptr, _stream, round_up


GRAD_BF16 = os.environ.get("HFC_GRAD_FMT", "fp16").strip().lower() == "bf16"
_FMT = int(GRAD_BF16)          # the `to_bf16` / `bf16` flags of the C ABI
_K_ROWS = 1 if GRAD_BF16 else 3    # hfc_im2col_t source kinds: fp32 rows -> bf16 | fp16
_K_ACT = 2 if GRAD_BF16 else 0     #                           fp16 act buffer -> bf16 | verbatim


# calibration only: device scalars max|operand| of every gradient operand written while a GradScale calibrates
_amax_log = None
# 1 / S of the GradScale whose backward is running: folded into every PARAMETER gradient where it is produced (the
# permute of the weight gradient, the bias column sums, the ChannelNorm parameter sums), so that parameter gradients are
# final -- and can be handed to a gradient all-reduce -- the moment their layer is done
_inv_scale = 1.0
# where finished parameter gradients go while a backward Function is still walking its layers (multi-GPU: the
# in-backward gradient reducer of hific_b200.dist); None = nowhere
_grad_sink = None


def emit(*tensors):
    """Called by the training plans when the gradients of a layer are final."""
    if _grad_sink is not None:
        _grad_sink.submit([t for t in tensors if t is not None])


def _log_operand(buf):
    if _amax_log is not None and not GRAD_BF16:
        _amax_log.append(buf.view(torch.float16).abs().amax().float())


class GradScale:
    """Power-of-two loss scale of ONE backward Function (fp16 gradient operands; a no-op with HFC_GRAD_FMT=bf16).

    `run(plan_backward, dout)` multiplies the incoming gradient by S, calls `plan_backward(scaled)` -> (dx, grads) with
    1 / S folded into every parameter gradient where it is produced (`_inv_scale`) and divides dx by S (all exact).  S is calibrated on the first call and then every HFC_GRAD_RECAL calls (default
    200) from the largest operand magnitude seen anywhere in that backward chain (one device->host read per
    calibration, none in between): it is placed near 2^PEAK_LOG2, i.e. 5 binades below fp16's saturation and 24
    above its smallest normal.  Between calibrations the conversions saturate, so a gradient that grew > 32 x clips."""
    PEAK_LOG2 = 11
    RECAL = int(os.environ.get("HFC_GRAD_RECAL", "200"))

    def __init__(self):
        self.scale, self.calls, self.peak = None, 0, None

    @staticmethod
    def _pow2_at_most(v):
        import math
        return 2.0 ** math.floor(math.log2(v))

    def run(self, plan_backward, dout):
        global _amax_log, _inv_scale, _grad_sink
        if GRAD_BF16:
            return plan_backward(dout)
        self.calls += 1

        def scaled(s):
            global _inv_scale
            _inv_scale = 1.0 / s
            try:
                return plan_backward(dout * s)
            finally:
                _inv_scale = 1.0
        if self.scale is None or (self.RECAL > 0 and self.calls % self.RECAL == 0):
            amax_in = float(dout.detach().abs().amax())
            if not (amax_in > 0.0) or amax_in == float("inf"):
                return plan_backward(dout)                  # all-zero (or broken) upstream gradient: nothing to place
            s = self.scale if self.scale is not None else self._pow2_at_most(1.0 / amax_in)
            res = None
            sink, _grad_sink = _grad_sink, None             # a calibration pass may repeat: its gradients are handed
            try:                                            # over in one piece afterwards
              for _ in range(4):
                _amax_log = []
                try:
                    res = scaled(s)
                    peak = float(torch.stack(_amax_log).max()) if _amax_log else 0.0
                finally:
                    _amax_log = None
                self.scale, self.peak = s, peak
                if peak <= 0.0:
                    break                                   # no 16-bit gradient operand on this path
                want = 2.0 ** self.PEAK_LOG2
                if peak < 65504.0 and want / 8 <= peak <= want * 4:
                    break
                # saturated: the true peak is unknown, step down far; else move the measured peak onto the target
                s = s / 4096.0 if peak >= 65504.0 else s * self._pow2_at_most(want / peak)
            finally:
                _grad_sink = sink
            emit(*res[1])
            s = self.scale
        else:
            s = self.scale
            res = scaled(s)
        dx, grads = res
        if dx is not None:
            dx = dx * (1.0 / s)
        return dx, grads


class Workspace:
    """Grow-only scratch buffers shared by all layers of a device (the backward pass is sequential)."""
    _bufs = {}

    @classmethod
    def get(cls, name, numel, dtype, device):
        key = (name, dtype, device.type, device.index)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < numel:
            buf = cls._bufs[key] = torch.empty(int(numel), dtype=dtype, device=device)
        return buf[:numel]

    @classmethod
    def named(cls, key, shape, dtype, device, zero=False):
        """A persistent buffer of its own (not shared between layers), optionally zero-initialised once."""
        k = (key, dtype, device.type, device.index)
        buf = cls._bufs.get(k)
        if buf is None:
            buf = cls._bufs[k] = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
        return buf


def _i8(vals):
    return bytes((v + 256) % 256 for v in vals)


def gemm_nt(a, a_bf16, b, b_bf16, m, n, k, out=None, k_splits=0):
    """C[m][n] fp32 = A[m][k] . B[n][k]^T for K-major 16-bit operands (raw buffers)."""
    ldc = round_up(n, 4)
    if out is None:
        out = torch.empty((m, ldc), dtype=torch.float32, device=a.device)
    check(lib.hfc_gemm_nt(_ptr(a), int(a_bf16), _ptr(b), int(b_bf16), m, n, k, _ptr(out), ldc, k_splits, _stream()),
          "gemm_nt")
    return out


def im2col_t(src, src_kind, n, hp, wp, cpad, c_src, gh, gw, stride, oh0, ow0, taps, c_rows, out):
    p_pad = out.shape[-1]
    dh, dw = _i8([t[0] for t in taps]), _i8([t[1] for t in taps])
    check(lib.hfc_im2col_t(_ptr(src), int(src_kind), n, hp, wp, cpad, c_src, gh, gw, stride, oh0, ow0, len(taps), dh, dw,
                           c_rows, p_pad, _ptr(out), _stream()), "im2col_t")
    return out


class ConvGrad:
    """Backward of the conv described by the same arguments as `ops.Conv` (forward geometry)."""

    def __init__(self, in_geom, cout, k, stride=1, transposed=False, pad_mode=PAD_ZERO, pad=(0, 0, 0, 0)):
        kh, kw = (k, k) if isinstance(k, int) else k
        assert kh == kw
        self.k, self.stride, self.transposed, self.pad_mode, self.pad = kh, stride, transposed, pad_mode, pad
        self.in_geom, self.cin, self.cout = in_geom, in_geom.c, cout
        n, h, w = in_geom.n, in_geom.h, in_geom.w
        pt, pl, pb, pr = pad
        if transposed:
            self.oh = (h - 1) * stride - 2 * pt + kh + (stride - 1)
            self.ow = (w - 1) * stride - 2 * pl + kw + (stride - 1)
        else:
            self.oh = (h + pt + pb - kh) // stride + 1
            self.ow = (w + pl + pr - kw) // stride + 1
        self.p_out, self.p_in = n * self.oh * self.ow, n * h * w
        self.dy_geom = Geom(n, self.oh, self.ow, cout, round_up(cout, 64))
        cin4 = round_up(self.cin, 4)
        self.fold = None
        self.dy_window_geom = None
        if not transposed and stride == 1 and cout <= 8 and kh <= 8 and kh > 1:
            # tiny cout (the 7x7 60 -> 3 head): padding the 3 gradient channels to a 64-channel K block would waste
            # 95 % of the MMA work, so dL/dy goes into an 8-channel-pitch buffer with a materialised ZERO border of
            # k-1 pixels and the conv runs in "window" packing (K = 8 pixels x 8 channels per filter row), exactly
            # like the forward 7x7 3 -> 60 encoder head.
            hq, wq = h + pt + pb, w + pl + pr
            b = kh - 1
            self.dy_window_geom = Geom(n, self.oh, self.ow, cout, 8, b, b, b, b + 1)
            self.dgrad = Conv(self.dy_window_geom, self.cin, kh, stride=1, pad_mode=PAD_REFLECT, pad=(b,) * 4, window=True,
                              out_mode=OUT_NHWC_F32, out_geom=Geom(n, hq, wq, self.cin, cin4), a_bf16=GRAD_BF16, b_bf16=GRAD_BF16, dgrad=True)
            self.fold = (hq, wq)
        elif not transposed and stride == 1:
            hq, wq = h + pt + pb, w + pl + pr
            self.dgrad = Conv(self.dy_geom, self.cin, kh, stride=1, pad_mode=PAD_ZERO, pad=(kh - 1,) * 4,
                              out_mode=OUT_NHWC_F32, out_geom=Geom(n, hq, wq, self.cin, cin4), a_bf16=GRAD_BF16, b_bf16=GRAD_BF16, dgrad=True)
            self.fold = (hq, wq)
        elif not transposed:
            hq = 2 * self.oh + kh - 1
            wq = 2 * self.ow + kw - 1
            self.dgrad = Conv(self.dy_geom, self.cin, kh, stride=2, transposed=True, pad=(0, 0, 0, 0),
                              out_mode=OUT_NHWC_F32, out_geom=Geom(n, hq, wq, self.cin, cin4), a_bf16=GRAD_BF16, b_bf16=GRAD_BF16)
            self.fold = (hq, wq)
        elif stride == 2:
            hi = kh - 2 - pt
            self.dgrad = Conv(self.dy_geom, self.cin, kh, stride=2, pad_mode=PAD_ZERO, pad=(pt, pl, hi, hi),
                              out_mode=OUT_NHWC_F32, out_geom=Geom(n, h, w, self.cin, cin4), a_bf16=GRAD_BF16, b_bf16=GRAD_BF16)
        else:
            self.dgrad = Conv(self.dy_geom, self.cin, kh, stride=1, pad_mode=PAD_ZERO, pad=(pt, pl, pt, pl),
                              out_mode=OUT_NHWC_F32, out_geom=Geom(n, h, w, self.cin, cin4), a_bf16=GRAD_BF16, b_bf16=GRAD_BF16)
        self.cin4 = cin4
        self.taps = [(ky, kx) for ky in range(kh) for kx in range(kw)]
        # weight-gradient GEMM operands: which side is transposed plainly (A1) and which is im2col'ed (A2).
        # The im2col side should be the one with FEWER channels (its matrix has ntaps times more rows).
        self.swap = (not transposed) and stride == 1 and cout <= 8
        # implicit weight gradient (csrc/wgrad_igemm.cu) for every layer whose operands are 64-channel-pitch NHWC
        # buffers; the 3-channel image layers (8-channel pitch input / tiny cout) keep the explicit transposed-im2col GEMM
        explicit = bool(os.environ.get("HFC_EXPLICIT_WGRAD"))
        pt, pl = pad[0], pad[1]
        k = kh
        d = _lib.WgradDesc()
        self.implicit = None
        if explicit:
            pass
        elif (not self.swap) and in_geom.cpad % 64 == 0:
            self.implicit = "plain"
            if transposed:       # P = x over the input pixels, S = dy sampled at i*s - p + k
                d.plain, d.shifted = in_geom.c_struct(), self.dy_geom.c_struct()
                self.wg_m, self.wg_c2 = self.cin, cout
            else:                # P = dy over the output pixels, S = x sampled at o*s + k - p (border-relative)
                d.plain, d.shifted = self.dy_geom.c_struct(), in_geom.c_struct()
                self.wg_m, self.wg_c2 = cout, self.cin
            offs = [(ky - pt, kx - pl) for ky, kx in self.taps]
            d.ntaps, d.stride, d.bf16, d.k_splits, d.window = len(self.taps), stride, _FMT, 0, 0
            self.wg_c2_rows = round_up(self.wg_c2, 64)
            self.wg_taps = self.taps
        elif (not self.swap) and (not transposed) and stride == 1 and in_geom.cpad == 8 and k <= 8:
            # 3-channel image layer (window-packed input): S rows are 8-pixel windows of x, the 8 column slots of a
            # filter row are kx = 0..7
            self.implicit = "window_x"
            d.plain, d.shifted = self.dy_geom.c_struct(), in_geom.c_struct()
            self.wg_m, self.wg_c2 = cout, self.cin
            offs = [(ky - pt, -pl) for ky in range(k)]
            d.ntaps, d.stride, d.bf16, d.k_splits, d.window = k, 1, _FMT, 0, 1
            self.wg_c2_rows = 8
            self.wg_taps = [(t // 8, t % 8) if t % 8 < k else (-1, -1) for t in range(8 * k)]
        elif self.swap and self.dy_window_geom is not None and pad_mode == PAD_REFLECT and \
                (in_geom.pt, in_geom.pl, in_geom.pb, in_geom.pr) == tuple(pad) and in_geom.cpad % 64 == 0:
            # tiny cout: P = x over its whole bordered buffer (taken as a border-less grid), S = 8-pixel windows of the
            # zero-bordered dy buffer; slot j of filter row ky holds kx = k-1-j
            self.implicit = "window_dy"
            hp, wp = h + in_geom.pt + in_geom.pb, w + in_geom.pl + in_geom.pr
            d.plain, d.shifted = Geom(n, hp, wp, self.cin, in_geom.cpad).c_struct(), self.dy_window_geom.c_struct()
            self.wg_m, self.wg_c2 = self.cin, cout
            offs = [(-ky, -(k - 1)) for ky in range(k)]
            d.ntaps, d.stride, d.bf16, d.k_splits, d.window = k, 1, _FMT, 0, 1
            self.wg_c2_rows = 8
            self.wg_taps = [(t // 8, k - 1 - t % 8) if t % 8 < k else (-1, -1) for t in range(8 * k)]
        if self.implicit:
            for i, (a, b) in enumerate(offs):
                d.tap_dh[i], d.tap_dw[i] = a, b
            self.wg_desc = d
            self.wg_cols = d.ntaps * (64 if d.window else self.wg_c2_rows)

    # ------------------------------------------------------------------ data gradient
    def data_grad(self, dy_rows, weight, out=None, scale=None, dy_act=None):
        """dy_rows: fp32 [n*oh*ow][ld >= cout] -> dx fp32 rows [n*h*w][cin4].  `scale`: optional device scalar
        multiplied into the weights while packing (1/sigma of a spectrally normalised layer).  dy_act: the same
        gradient already in operand format (then dy_rows may be None)."""
        sk = dict(scale=scale, scale_key=object()) if scale is not None else {}
        dev = (dy_rows if dy_rows is not None else dy_act).device
        if dy_act is None:
            dy_act = self.dy_to_act(dy_rows)
        if self.fold is None:
            if out is None:
                out = torch.empty((self.p_in, self.cin4), dtype=torch.float32, device=dev)
            self.dgrad(dy_act.view(torch.float16), weight, out=out, **sk)
            return out
        hq, wq = self.fold
        dxp = Workspace.get("dxp", self.in_geom.n * hq * wq * self.cin4, torch.float32, dev).view(-1, self.cin4)
        self.dgrad(dy_act.view(torch.float16), weight, out=dxp, **sk)
        if out is None:
            out = torch.empty((self.p_in, self.cin4), dtype=torch.float32, device=dev)
        n, h, w = self.in_geom.n, self.in_geom.h, self.in_geom.w
        pt, pl, pb, pr = self.pad
        fg = Geom(n, h, w, self.cin4, self.cin4, pt, pl, pb, pr).c_struct()
        check(lib.hfc_pad_fold(_ptr(dxp), self.cin4, hq, wq, ctypes.byref(fg), int(self.pad_mode == PAD_REFLECT),
                               _ptr(out), self.cin4, _stream()), "pad_fold")
        return out

    # ------------------------------------------------------------------ weight / bias gradient
    def dy_to_act(self, dy_rows):
        """fp32 gradient rows -> NHWC bf16 buffer (the operand format of the backward GEMMs): border-less with a
        64-channel pitch, or -- tiny cout -- 8-channel pitch with a zero border of k-1 pixels (window packing)."""
        if self.dy_window_geom is not None:
            g = self.dy_window_geom
            buf = Workspace.named(("dy_window",) + g.shape, g.shape, torch.int16, dy_rows.device, zero=True)
            gs = g.c_struct()      # the border is zeroed once; only the interior is rewritten
            check(lib.hfc_rows_to_act_geom(_ptr(dy_rows), dy_rows.shape[-1], ctypes.byref(gs), _FMT, _ptr(buf), _stream()),
                  "rows_to_act_geom")
            _log_operand(buf)
            return buf
        g = self.dy_geom
        dy_act = Workspace.get("dy_act", g.n * g.h * g.w * g.cpad, torch.int16, dy_rows.device).view(g.shape)
        check(lib.hfc_rows_to_act(_ptr(dy_rows), dy_rows.shape[-1], self.p_out, self.cout, g.cpad, _FMT, _ptr(dy_act),
                                  _stream()), "rows_to_act")
        _log_operand(dy_act)
        return dy_act

    def _weight_grad_implicit(self, x_act, dy_rows, dw_out, accumulate, scale, dy_act=None):
        dev = x_act.device
        if dy_act is None:
            dy_act = self.dy_to_act(dy_rows)
        if GRAD_BF16:
            x_bf = Workspace.get("x_bf16", x_act.numel(), torch.int16, dev)
            check(lib.hfc_act_to_bf16(_ptr(x_act), _ptr(x_bf), x_act.numel(), _stream()), "act_to_bf16")
        else:
            x_bf = x_act              # fp16 x fp16: the saved activation buffer is the operand
        ncols = self.wg_cols
        cbuf = Workspace.get("wgrad_c", self.wg_m * ncols, torch.float32, dev).view(self.wg_m, ncols)
        x_is_plain = self.transposed or self.implicit == "window_dy"
        plain, shifted = (x_bf, dy_act) if x_is_plain else (dy_act, x_bf)
        check(lib.hfc_wgrad(ctypes.byref(self.wg_desc), _ptr(plain), _ptr(shifted), _ptr(cbuf), ncols, _stream()), "wgrad")
        k = self.k
        ky, kx = _i8([t[0] for t in self.wg_taps]), _i8([t[1] for t in self.wg_taps])
        shape = (self.cin, self.cout, k, k) if self.transposed else (self.cout, self.cin, k, k)
        if dw_out is None:
            dw_out = torch.empty(shape, dtype=torch.float32, device=dev)
        if self.implicit == "window_dy":
            # C is [ci][(slot, co)] but dW is [co][ci][ky][kx]: permute into a temporary and transpose the leading dims
            tmp = torch.empty((self.cin, self.cout, k, k), dtype=torch.float32, device=dev)
            check(lib.hfc_permute_wgrad(_ptr(cbuf), ncols, self.wg_m, self.wg_c2, self.wg_c2_rows, k, k, len(self.wg_taps),
                                        ky, kx, float(scale) * _inv_scale, 0, _ptr(tmp), _stream()), "permute_wgrad")
            if accumulate:
                dw_out += tmp.transpose(0, 1)
            else:
                dw_out.copy_(tmp.transpose(0, 1))
            return dw_out
        check(lib.hfc_permute_wgrad(_ptr(cbuf), ncols, self.wg_m, self.wg_c2, self.wg_c2_rows, k, k, len(self.wg_taps), ky, kx,
                                    float(scale) * _inv_scale, int(accumulate), _ptr(dw_out), _stream()), "permute_wgrad")
        return dw_out

    def weight_grad(self, x_act, dy_rows, dw_out=None, accumulate=False, scale=1.0, dy_act=None):
        """x_act: the forward input act buffer (fp16, with its border); dy_rows fp32 [n*oh*ow][ld].
        Returns dW in the torch layout of the forward weight."""
        if self.implicit:
            return self._weight_grad_implicit(x_act, dy_rows, dw_out, accumulate, scale, dy_act)
        assert dy_rows is not None, "the explicit weight-gradient path needs the fp32 gradient rows"
        dev = dy_rows.device
        ig = self.in_geom
        n, h, w = ig.n, ig.h, ig.w
        hp, wp = h + ig.pt + ig.pb, w + ig.pl + ig.pr
        ld_dy = dy_rows.shape[-1]
        pt, pl = self.pad[0], self.pad[1]
        k, s = self.k, self.stride
        if not self.transposed and not self.swap:
            # pixels = output pixels; A1 = dy^T (bf16), COLT = im2col(x)^T (fp16)
            p_pad = round_up(self.p_out, 64)
            c1_rows, c2_rows = round_up(self.cout, 64), (8 if ig.cpad == 8 else round_up(self.cin, 64))
            a1 = Workspace.get("a1t", c1_rows * p_pad, torch.int16, dev).view(c1_rows, p_pad)
            im2col_t(dy_rows, _K_ROWS, n, self.oh, self.ow, ld_dy, self.cout, self.oh, self.ow, 1, 0, 0, [(0, 0)], c1_rows, a1)
            col = Workspace.get("colt", len(self.taps) * c2_rows * p_pad, torch.int16, dev).view(-1, p_pad)
            im2col_t(x_act, _K_ACT, n, hp, wp, ig.cpad, self.cin, self.oh, self.ow, s, ig.pt - pt, ig.pl - pl, self.taps,
                     c2_rows, col)
            m, c2, a_bf, b_bf = self.cout, self.cin, GRAD_BF16, GRAD_BF16
            shape = (self.cout, self.cin, k, k)
        elif not self.transposed:
            # tiny cout: pixels = padded input pixels q; A1 = x^T (fp16), COLT[(tap, co)][q] = dy[q - tap] (bf16)
            if self.pad_mode == PAD_REFLECT:
                assert (ig.pt, ig.pl, ig.pb, ig.pr) == tuple(self.pad), "swap wgrad expects the border to equal the padding"
                gh, gw, o0h, o0w = hp, wp, 0, 0
            else:
                gh, gw, o0h, o0w = h + self.pad[0] + self.pad[2], w + self.pad[1] + self.pad[3], -pt, -pl
            p_pad = round_up(n * gh * gw, 64)
            c1_rows, c2_rows = round_up(self.cin, 64), 8
            a1 = Workspace.get("a1t", c1_rows * p_pad, torch.int16, dev).view(c1_rows, p_pad)
            im2col_t(x_act, _K_ACT, n, hp, wp, ig.cpad, self.cin, gh, gw, 1, o0h, o0w, [(0, 0)], c1_rows, a1)
            col = Workspace.get("colt", len(self.taps) * c2_rows * p_pad, torch.int16, dev).view(-1, p_pad)
            im2col_t(dy_rows, _K_ROWS, n, self.oh, self.ow, ld_dy, self.cout, gh, gw, 1, 0, 0,
                     [(-ky, -kx) for ky, kx in self.taps], c2_rows, col)
            m, c2, a_bf, b_bf = self.cin, self.cout, GRAD_BF16, GRAD_BF16
            shape = (self.cout, self.cin, k, k)
        else:
            # transposed conv: pixels = input pixels i; A1 = x^T (fp16), COLT[(tap, co)][i] = dy[i*s - p + tap] (bf16)
            p_pad = round_up(self.p_in, 64)
            c1_rows, c2_rows = round_up(self.cin, 64), round_up(self.cout, 64)
            a1 = Workspace.get("a1t", c1_rows * p_pad, torch.int16, dev).view(c1_rows, p_pad)
            im2col_t(x_act, _K_ACT, n, hp, wp, ig.cpad, self.cin, h, w, 1, ig.pt, ig.pl, [(0, 0)], c1_rows, a1)
            col = Workspace.get("colt", len(self.taps) * c2_rows * p_pad, torch.int16, dev).view(-1, p_pad)
            im2col_t(dy_rows, _K_ROWS, n, self.oh, self.ow, ld_dy, self.cout, h, w, s, -pt, -pl, self.taps, c2_rows, col)
            m, c2, a_bf, b_bf = self.cin, self.cout, GRAD_BF16, GRAD_BF16
            shape = (self.cin, self.cout, k, k)
        ncols = len(self.taps) * c2_rows
        cbuf = Workspace.get("wgrad_c", m * round_up(ncols, 4), torch.float32, dev).view(m, round_up(ncols, 4))
        gemm_nt(a1, a_bf, col, b_bf, m, ncols, p_pad, out=cbuf)
        if dw_out is None:
            dw_out = torch.empty(shape, dtype=torch.float32, device=dev)
        ky, kx = _i8([t[0] for t in self.taps]), _i8([t[1] for t in self.taps])
        if self.swap:
            # C is [ci][(tap, co)] but dW is [co][ci][ky][kx]: permute into a temporary and transpose the two leading dims
            tmp = torch.empty((self.cin, self.cout, k, k), dtype=torch.float32, device=dev)
            check(lib.hfc_permute_wgrad(_ptr(cbuf), cbuf.shape[1], m, c2, c2_rows, k, k, len(self.taps), ky, kx,
                                        float(scale) * _inv_scale, 0, _ptr(tmp), _stream()), "permute_wgrad")
            if accumulate:
                dw_out += tmp.transpose(0, 1)
            else:
                dw_out.copy_(tmp.transpose(0, 1))
            return dw_out
        check(lib.hfc_permute_wgrad(_ptr(cbuf), cbuf.shape[1], m, c2, c2_rows, k, k, len(self.taps), ky, kx,
                                    float(scale) * _inv_scale, int(accumulate), _ptr(dw_out), _stream()), "permute_wgrad")
        return dw_out

    def bias_grad(self, dy_rows, db_out=None, accumulate=False, scale=1.0):
        if db_out is None:
            db_out = torch.zeros(self.cout, dtype=torch.float32, device=dy_rows.device)
        elif not accumulate:
            db_out.zero_()
        check(lib.hfc_col_sums(_ptr(dy_rows), dy_rows.shape[-1], self.p_out, self.cout, float(scale) * _inv_scale, _ptr(db_out),
                               _stream()), "col_sums")
        return db_out
