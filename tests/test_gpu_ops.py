"""GPU unit tests of the individual libhfc kernels against plain torch fp32 on the same device.

Inputs and weights are pre-rounded to fp16 so that the only difference between the tcgen05 path
(fp16 operands, fp32 accumulate) and torch fp32 (TF32 disabled) is summation order: tolerances are
therefore tight (1e-4 relative for fp32 outputs, fp16 resolution for fp16 outputs).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # collected on the CPU box, skipped there
    pytest.skip("needs a CUDA device", allow_module_level=True)

from hific_b200 import ops  # noqa: E402
from hific_b200.ops import (ACT_LEAKY02, ACT_NONE, ACT_RELU, OUT_NCHW_F32, OUT_NHWC_F16, OUT_NHWC_F32,  # noqa: E402
                            PAD_REFLECT, PAD_ZERO, Conv, Geom, round_up)

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"


def r16(t):
    return t.half().float()


def channel_norm_ref(x, gamma, beta, eps=1e-3):
    mu = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, keepdim=True)
    return gamma.view(1, -1, 1, 1) * ((x - mu) * torch.rsqrt(var + eps)) + beta.view(1, -1, 1, 1)


def act_ref(x, act):
    if act == ACT_RELU:
        return F.relu(x)
    if act == ACT_LEAKY02:
        return F.leaky_relu(x, 0.2)
    return x


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def run_conv_case(n, cin, h, w, cout, k, stride=1, pad=(0, 0, 0, 0), pad_mode=PAD_ZERO, transposed=False,
                  out_mode=OUT_NCHW_F32, out_border=(0, 0, 0, 0), act=ACT_NONE, norm=False, window=False,
                  block_n=0, seed=0, cluster=(0, 0), wide=0, pair=0, expect=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    kh, kw = (k, k) if isinstance(k, int) else k
    x = r16(torch.randn(n, cin, h, w, generator=g)).to(DEV)
    if transposed:
        wt = r16(torch.randn(cin, cout, kh, kw, generator=g) / math.sqrt(cin * kh * kw)).to(DEV)
    else:
        wt = r16(torch.randn(cout, cin, kh, kw, generator=g) / math.sqrt(cin * kh * kw)).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    gamma = (1 + 0.1 * torch.randn(cout, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(cout, generator=g)).to(DEV)

    pt, pl, pb, pr = pad
    cpad_in = 8 if window else round_up(cin, 64)
    if pad_mode == PAD_REFLECT:
        in_border = (pt, pl, pb, pr + (1 if window else 0))
        in_geom = Geom(n, h, w, cin, cpad_in, *in_border)
    else:
        in_geom = Geom(n, h, w, cin, cpad_in)
    x_act = ops.nchw_to_act(x, in_geom, reflect=(pad_mode == PAD_REFLECT))

    # reference
    if transposed:
        ref = F.conv_transpose2d(x, wt, bias, stride=stride, padding=pt, output_padding=stride - 1)
    else:
        xp = F.pad(x, (pl, pr, pt, pb), mode="reflect" if pad_mode == PAD_REFLECT else "constant")
        ref = F.conv2d(xp, wt, bias, stride=stride)
    if norm:
        ref = channel_norm_ref(ref, gamma, beta)
    ref = act_ref(ref, act)
    oh, ow = ref.shape[2:]

    if out_mode == OUT_NHWC_F16:
        out_geom = Geom(n, oh, ow, cout, round_up(cout, 64), *out_border)
    elif out_mode == OUT_NHWC_F32:
        out_geom = Geom(n, oh, ow, cout, round_up(cout, 4))
    else:
        out_geom = Geom(n, oh, ow, cout, cout)
    conv = Conv(in_geom, cout, k, stride=stride, transposed=transposed, pad_mode=pad_mode, pad=pad,
                out_mode=out_mode, out_geom=out_geom, out_reflect=any(out_border), act=act, norm=norm,
                window=window, block_n=block_n, cluster=cluster, wide=wide, pair=pair)
    if expect is not None:
        for k, v in expect.items():
            assert getattr(conv.info, k) == v, (k, getattr(conv.info, k), v)
    out = conv(x_act, wt, bias, gamma if norm else None, beta if norm else None)
    torch.cuda.synchronize()

    if out_mode == OUT_NCHW_F32:
        err = rel_err(out, ref)
        assert err < 2e-5, f"NCHW rel err {err}"
        assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4)
    elif out_mode == OUT_NHWC_F32:
        got = out.view(n, oh, ow, -1)[..., :cout].permute(0, 3, 1, 2)
        err = rel_err(got, ref)
        assert err < 2e-5, f"NHWC32 rel err {err}"
        if out_geom.cpad > cout:
            assert out.view(n, oh, ow, -1)[..., cout:].abs().max().item() == 0.0
    else:
        bt, bl, bb, br = out_border
        full = F.pad(ref, (bl, br, bt, bb), mode="reflect") if any(out_border) else ref
        got = out[..., :cout].permute(0, 3, 1, 2).float()
        assert got.shape == full.shape, (got.shape, full.shape)
        assert torch.allclose(got, full, rtol=2e-3, atol=2e-3), f"max err {(got - full).abs().max().item()}"
        if out_geom.cpad > cout:
            assert out[..., cout:].float().abs().max().item() == 0.0
    return conv


def test_conv_basic_nchw():
    run_conv_case(2, 64, 16, 16, 64, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT)


def test_conv_two_chunks_nhwc16_border():
    run_conv_case(2, 128, 16, 16, 240, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 1, 1, 1), act=ACT_RELU)


def test_conv_channel_padding_60():
    run_conv_case(2, 60, 16, 16, 120, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 0, 0, 1), norm=True, act=ACT_RELU)


def test_conv_stride2_asym_norm():
    run_conv_case(2, 64, 32, 32, 128, 3, stride=2, pad=(1, 0, 0, 1), pad_mode=PAD_REFLECT,
                  out_mode=OUT_NHWC_F16, out_border=(1, 0, 0, 1), norm=True, act=ACT_RELU)


def test_conv_window_7x7_first_layer():
    run_conv_case(2, 3, 32, 32, 60, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 0, 0, 1), norm=True, act=ACT_RELU, window=True)


@pytest.mark.parametrize("w", [128, 200, 256])
def test_conv_window_7x7_wide_images_flat_segment(w):
    """Images >= 96 pixels wide: the window operand is served from a plain pixel segment through an un-swizzled UMMA
    descriptor with overlapping rows (tile = 128 pixels of one row; ragged last tile at w = 200)."""
    run_conv_case(2, 3, 12, w, 60, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 0, 0, 1), norm=True, act=ACT_RELU, window=True)


def test_conv_multi_ntile_nhwc32():
    run_conv_case(2, 64, 16, 16, 320, 3, pad=(1, 1, 1, 1), pad_mode=PAD_ZERO, out_mode=OUT_NHWC_F32, act=ACT_RELU)


def test_conv_5x5_s2_small_maps():
    run_conv_case(3, 320, 8, 8, 320, 5, stride=2, pad=(2, 2, 2, 2), pad_mode=PAD_REFLECT, out_mode=OUT_NCHW_F32)


def test_conv_5x5_s2_to_nhwc16_border2():
    run_conv_case(3, 320, 16, 16, 320, 5, stride=2, pad=(2, 2, 2, 2), pad_mode=PAD_REFLECT,
                  out_mode=OUT_NHWC_F16, out_border=(2, 2, 2, 2), act=ACT_RELU)


def test_conv_cout_220_block224():
    run_conv_case(2, 128, 16, 16, 220, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16)


def test_conv_head_7x7_cout3():
    run_conv_case(1, 60, 32, 32, 3, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT)


def test_conv_leaky_cout12():
    run_conv_case(2, 220, 16, 16, 12, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  act=ACT_LEAKY02)


def test_convT_k3_s2_norm_border3():
    run_conv_case(2, 128, 8, 8, 60, 3, stride=2, pad=(1, 1, 1, 1), transposed=True, out_mode=OUT_NHWC_F16,
                  out_border=(3, 3, 3, 4), norm=True, act=ACT_RELU)


def test_convT_k5_s2():
    run_conv_case(3, 320, 4, 4, 320, 5, stride=2, pad=(2, 2, 2, 2), transposed=True, out_mode=OUT_NHWC_F16,
                  act=ACT_RELU)


def test_convT_k3_s1_nchw():
    run_conv_case(2, 320, 16, 16, 220, 3, stride=1, pad=(1, 1, 1, 1), transposed=True)


def test_convT_k3_s2_nhwc32():
    run_conv_case(2, 192, 16, 16, 480, 3, stride=2, pad=(1, 1, 1, 1), transposed=True, out_mode=OUT_NHWC_F32)


def test_conv_ragged_sizes():
    run_conv_case(3, 64, 19, 27, 96, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 1, 1, 1))
    run_conv_case(1, 64, 22, 38, 64, 3, stride=2, pad=(1, 0, 0, 1), pad_mode=PAD_REFLECT)


def test_conv_resblock_shape_many_tiles():
    # 960 -> 960, 16x16, batch 4: 4 M tiles x 4 N tiles, K = 8640 (135 k-blocks, pipeline wraps many times)
    run_conv_case(4, 960, 16, 16, 960, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F32)


def test_conv_persistent_many_tiles():
    # more tiles than SMs so every CTA loops (TMEM double buffering exercised)
    run_conv_case(8, 64, 64, 64, 64, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 1, 1, 1), norm=True, act=ACT_RELU)


def test_weight_repack_on_update():
    conv = run_conv_case(2, 64, 16, 16, 64, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT)
    w = torch.nn.Parameter(r16(torch.randn(64, 64, 3, 3, device=DEV) * 0.05))
    p1 = conv.packed_weights(w).clone()
    with torch.no_grad():
        w.mul_(2.0)
    p2 = conv.packed_weights(w)
    assert torch.allclose(p2.float(), 2 * p1.float())


def test_nchw_to_act_norm_and_border():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 220, 16, 16, generator=g).to(DEV)
    gamma = (1 + 0.1 * torch.randn(220, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(220, generator=g)).to(DEV)
    geom = Geom(2, 16, 16, 220, 256, 1, 1, 1, 1)
    buf = ops.nchw_to_act(x, geom, reflect=True, norm=True, gamma=gamma, beta=beta)
    ref = F.pad(channel_norm_ref(x, gamma, beta), (1, 1, 1, 1), mode="reflect")
    got = buf[..., :220].permute(0, 3, 1, 2).float()
    assert torch.allclose(got, ref, rtol=2e-3, atol=2e-3)
    assert buf[..., 220:].float().abs().max().item() == 0.0


def test_channelnorm_standalone_residuals():
    g = torch.Generator().manual_seed(4)
    n, c, h, w = 2, 960, 16, 16
    x = torch.randn(n, c, h, w, generator=g).to(DEV) * 3 + 0.5
    r1 = torch.randn(n, c, h, w, generator=g).to(DEV)
    r2 = torch.randn(n, c, h, w, generator=g).to(DEV)
    gamma = (1 + 0.1 * torch.randn(c, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(c, generator=g)).to(DEV)
    rows = lambda t: t.permute(0, 2, 3, 1).reshape(-1, c).contiguous()
    geom = Geom(n, h, w, c, 960, 1, 1, 1, 1)
    out_act, out_f32 = ops.channelnorm(rows(x), geom, gamma, beta, act=ACT_NONE, reflect=True, res1=rows(r1),
                                       res2=rows(r2), want_f32=True)
    ref = channel_norm_ref(x, gamma, beta) + r1 + r2
    assert rel_err(out_f32, rows(ref)) < 1e-5
    got = out_act.permute(0, 3, 1, 2).float()
    assert torch.allclose(got, F.pad(ref, (1, 1, 1, 1), mode="reflect"), rtol=2e-3, atol=2e-3)
    # relu, 480 channels padded to 512, no residual
    c = 480
    x = torch.randn(n, c, h, w, generator=g).to(DEV)
    gamma = torch.ones(c, device=DEV)
    beta = torch.zeros(c, device=DEV)
    geom = Geom(n, h, w, c, 512)
    out_act, _ = ops.channelnorm(x.permute(0, 2, 3, 1).reshape(-1, c).contiguous(), geom, gamma, beta, act=ACT_RELU)
    ref = F.relu(channel_norm_ref(x, gamma, beta))
    assert torch.allclose(out_act[..., :c].permute(0, 3, 1, 2).float(), ref, rtol=2e-3, atol=2e-3)
    assert out_act[..., c:].float().abs().max().item() == 0.0


def _latent_ref(y, mu, sraw, noise, lb, kind):
    sc = torch.clamp(sraw, min=lb)
    cdf = (lambda v: 0.5 * torch.erfc(v * (-1.0 / math.sqrt(2)))) if kind == "gaussian" else torch.sigmoid

    def lik(x):
        d = (x - mu).abs()
        p = cdf((0.5 - d) / sc) - cdf(-(0.5 + d) / sc)
        return torch.clamp(p, min=1e-9)

    yq = torch.floor(y - mu + 0.5) + mu
    sq = torch.log(lik(yq) + 1e-9).double().sum()
    sn = torch.log(lik(y + noise) + 1e-9).double().sum()
    return yq, sn, sq


@pytest.mark.parametrize("kind", ["gaussian", "logistic"])
def test_latent_likelihood(kind):
    g = torch.Generator().manual_seed(5)
    shape = (4, 220, 16, 16)
    y = (2 * torch.randn(shape, generator=g)).to(DEV)
    mu = torch.randn(shape, generator=g).to(DEV)
    sraw = (2 * torch.rand(shape, generator=g)).to(DEV)
    noise = (torch.rand(shape, generator=g) - 0.5).to(DEV)
    dec, sums = ops.latent_likelihood(y, mu, sraw, noise, 0.11, kind)
    yq, sn, sq = _latent_ref(y, mu, sraw, noise, 0.11, kind)
    assert torch.allclose(dec, yq, rtol=0, atol=1e-5)
    assert abs(sums[0].item() - sn.item()) / abs(sn.item()) < 2e-5
    assert abs(sums[1].item() - sq.item()) / abs(sq.item()) < 2e-5


def test_latent_likelihood_ragged_count():
    g = torch.Generator().manual_seed(6)
    shape = (1, 3, 5, 7)  # 105 elements: vector body + scalar tail
    y = torch.randn(shape, generator=g).to(DEV)
    mu = torch.randn(shape, generator=g).to(DEV)
    sraw = torch.rand(shape, generator=g).to(DEV)
    noise = (torch.rand(shape, generator=g) - 0.5).to(DEV)
    dec, sums = ops.latent_likelihood(y, mu, sraw, noise)
    yq, sn, sq = _latent_ref(y, mu, sraw, noise, 0.11, "gaussian")
    assert torch.allclose(dec, yq, atol=1e-5)
    assert abs(sums[0].item() - sn.item()) < 1e-3 and abs(sums[1].item() - sq.item()) < 1e-3


def test_hyperlatent_likelihood():
    g = torch.Generator().manual_seed(7)
    C = 320
    filters = (1, 3, 3, 3, 1)
    Hs, a_s, bs = [], [], []
    for k in range(4):
        Hs.append(torch.randn(C, filters[k + 1], filters[k], generator=g).to(DEV))
        a_s.append((0.5 * torch.randn(C, filters[k + 1], 1, generator=g)).to(DEV))
        bs.append((torch.rand(C, filters[k + 1], 1, generator=g) - 0.5).to(DEV))
    z = (3 * torch.randn(4, C, 4, 4, generator=g)).to(DEV)
    noise = (torch.rand(z.shape, generator=g) - 0.5).to(DEV)

    def cdf_logits(x):
        logits = x
        for k in range(4):
            logits = torch.bmm(F.softplus(Hs[k]), logits) + bs[k]
            logits = logits + torch.tanh(a_s[k]) * torch.tanh(logits)
        return logits

    def loglik_sum(x):
        lat = x.permute(1, 0, 2, 3).reshape(C, 1, -1)
        u, l = cdf_logits(lat + 0.5), cdf_logits(lat - 0.5)
        s = -torch.sign(u + l)
        p = (torch.sigmoid(s * u) - torch.sigmoid(s * l)).abs().clamp(min=1e-9)
        return torch.log(p + 1e-9).double().sum()

    params = ops.pack_density_params(Hs, a_s, bs)
    zn, zq, sums = ops.hyperlatent_likelihood(z, params, noise)
    assert torch.equal(zq, torch.floor(z + 0.5))
    assert torch.equal(zn, z + noise)
    rn, rq = loglik_sum(z + noise), loglik_sum(torch.floor(z + 0.5))
    assert abs(sums[0].item() - rn.item()) / abs(rn.item()) < 2e-5
    assert abs(sums[1].item() - rq.item()) / abs(rq.item()) < 2e-5


def test_adam_matches_torch():
    """hific_b200.optim.Adam (one launch per step) against torch.optim.Adam over several steps, odd sizes included."""
    from hific_b200.optim import Adam
    g = torch.Generator().manual_seed(5)
    shapes = [(960, 960, 3, 3), (7,), (1, 60, 1, 1), (33, 5), (8192 * 3 + 1,)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    for wd in (0.0, 0.01):
        oa, ob = Adam(ours, lr=1e-3, weight_decay=wd), torch.optim.Adam(ref, lr=1e-3, weight_decay=wd)
        for step in range(4):
            for p, q in zip(ours, ref):
                grad = torch.randn(p.shape, generator=g).to(DEV) * (10.0 ** (step - 2))
                p.grad, q.grad = grad.clone(), grad.clone()
            if step == 2:
                ours[1].grad = ref[1].grad = None          # parameters without a gradient are skipped
            l0 = ops.launch_count()
            ver = ours[0]._version
            oa.step(); ob.step()
            assert ours[0]._version > ver       # the packed-weight caches key on the version counter
            assert ops.launch_count() - l0 == (2 if step == 3 else 1)   # one launch per distinct step count
        for p, q in zip(ours, ref):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
        sa, sb = oa.state_dict()["state"], ob.state_dict()["state"]
        assert torch.allclose(sa[0]["exp_avg_sq"], sb[0]["exp_avg_sq"], rtol=1e-5, atol=1e-12)
