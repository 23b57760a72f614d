"""GPU tests of the two traffic-saving modes of the conv kernel: thread-block clusters with TMA multicast of
the shared operand tiles, and the row-resident 'wide' mode (halo row + resident weights + descriptor shift).
Kept in their own file so that a protocol bug here (trap) cannot poison the CUDA context of the other tests."""
import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from test_gpu_ops import run_conv_case  # noqa: E402
from hific_b200.ops import (ACT_RELU, OUT_NCHW_F32, OUT_NHWC_F16, OUT_NHWC_F32, PAD_REFLECT, PAD_ZERO)  # noqa: E402


@pytest.mark.parametrize("cluster", [(2, 1), (1, 2), (2, 2)])
def test_cluster_multicast_resblock_shape(cluster):
    run_conv_case(4, 960, 16, 16, 960, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F32,
                  cluster=cluster, expect=dict(cluster_m=cluster[0], cluster_n=cluster[1]))


def test_cluster_auto_on_big_layer():
    # 32 x 32x32 -> 256 M tiles, 1 N tile: auto picks 2x1 (weights shared by two pixel tiles).  (Maps of >= 4 096 pixels
    # per image with one N tile of <= 128 columns take the thin instantiation instead: single CTAs, no cluster.)
    run_conv_case(32, 64, 32, 32, 120, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 0, 0, 1), norm=True, act=ACT_RELU, expect=dict(cluster_m=2, cluster_n=1))
    run_conv_case(8, 64, 64, 64, 120, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 0, 0, 1), norm=True, act=ACT_RELU, expect=dict(cluster_m=1, cluster_n=1))


def test_cluster_with_padding_tiles():
    # odd tile counts in both directions: 3 M tiles (batch 3 of 16x8), 3 N tiles of 160 -> dummy tiles
    run_conv_case(3, 64, 8, 16, 480, 3, pad=(1, 1, 1, 1), pad_mode=PAD_ZERO, out_mode=OUT_NHWC_F32, block_n=160,
                  cluster=(2, 2))


def test_cluster_stride2_and_small_maps():
    run_conv_case(16, 128, 32, 32, 240, 3, stride=2, pad=(1, 0, 0, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 1, 1, 1), norm=True, act=ACT_RELU, cluster=(2, 1))
    # 4x4 maps: tile = 4x4x8 images, A slice split along the batch dimension of the box
    run_conv_case(32, 320, 8, 8, 320, 5, stride=2, pad=(2, 2, 2, 2), pad_mode=PAD_REFLECT, out_mode=OUT_NCHW_F32,
                  cluster=(1, 2), block_n=160)


def test_cluster_transposed():
    run_conv_case(8, 256, 32, 32, 120, 3, stride=2, pad=(1, 1, 1, 1), transposed=True, out_mode=OUT_NHWC_F16,
                  norm=True, act=ACT_RELU, cluster=(2, 1))


@pytest.mark.parametrize("w", [128, 256, 200])
def test_wide_mode_7x7_head(w):
    run_conv_case(2, 60, 24, w, 3, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT, wide=1, expect=dict(wide=1))


def test_wide_mode_forced_3x3_and_forbidden():
    run_conv_case(2, 64, 16, 128, 16, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, wide=1, expect=dict(wide=1))
    run_conv_case(1, 60, 32, 128, 3, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT, wide=2, expect=dict(wide=0))


@pytest.mark.parametrize("cluster", [(2, 1), (2, 2)])
def test_cta_pair_resblock_shape(cluster):
    run_conv_case(4, 960, 16, 16, 960, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F32,
                  cluster=cluster, pair=1, expect=dict(cluster_m=2, cluster_n=cluster[1], pair=1))


def test_cta_pair_many_tiles_persistent():
    # 16 x 32x32 -> 128 M tiles x 2 N tiles (block_n 240): every pair loops over several tiles
    run_conv_case(16, 256, 32, 32, 480, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F32,
                  cluster=(2, 2), pair=1, expect=dict(pair=1))


def test_cta_pair_stride2_fused_norm_and_padding_tiles():
    run_conv_case(16, 128, 32, 32, 240, 3, stride=2, pad=(1, 0, 0, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 1, 1, 1), norm=True, act=ACT_RELU, cluster=(2, 1), pair=1, expect=dict(pair=1))
    # odd number of M tiles (3) and of N tiles (3): dummy tiles inside the pair / cluster
    run_conv_case(3, 64, 8, 16, 480, 3, pad=(1, 1, 1, 1), pad_mode=PAD_ZERO, out_mode=OUT_NHWC_F32, block_n=160,
                  cluster=(2, 2), pair=1, expect=dict(pair=1))


def test_cta_pair_transposed_and_small_maps():
    run_conv_case(8, 256, 32, 32, 120, 3, stride=2, pad=(1, 1, 1, 1), transposed=True, out_mode=OUT_NHWC_F16,
                  norm=True, act=ACT_RELU, cluster=(2, 1), pair=1, expect=dict(pair=1))
    run_conv_case(32, 320, 8, 8, 320, 5, stride=2, pad=(2, 2, 2, 2), pad_mode=PAD_REFLECT, out_mode=OUT_NCHW_F32,
                  cluster=(2, 2), block_n=160, pair=1, expect=dict(pair=1))


@pytest.mark.parametrize("w", [128, 256, 200, 122, 123])
def test_tap_in_n_mode_7x7_head(w):
    run_conv_case(2, 60, 20, w, 3, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT, expect=dict(tapn=1))


def test_tap_in_n_other_shapes():
    run_conv_case(1, 64, 16, 128, 4, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, expect=dict(tapn=1))       # 3x3, cout 4
    run_conv_case(2, 128, 9, 130, 2, 5, pad=(2, 2, 2, 2), pad_mode=PAD_ZERO, expect=dict(tapn=1))          # 2 K chunks, zero pad
    run_conv_case(1, 60, 16, 128, 3, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT, wide=1, expect=dict(tapn=0, wide=1))


def test_cta_pair_two_n_tiles_per_item():
    """Bench-size residual conv (batch 32: 64 M tiles x 4 N tiles): the planner gives every CTA pair two N tiles (one
    per TMEM half) against one fetch of the A tile; same numbers as the single-tile schedule."""
    run_conv_case(32, 960, 16, 16, 960, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F32,
                  expect=dict(pair=1, nsub=2))
    run_conv_case(32, 960, 16, 16, 960, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F16,
                  out_border=(1, 1, 1, 1), act=ACT_RELU, expect=dict(pair=1, nsub=2))
    # odd number of N-tile pairs is impossible (n_tiles must be even); 480 = 2 x 240 -> one group
    run_conv_case(32, 960, 16, 16, 480, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT, out_mode=OUT_NHWC_F32,
                  expect=dict(pair=1))


# ----------------------------------------------------------------------------------------------------------------------
# Non-default generator variants on the kernels: sample_noise=True (992-channel trunk: 1 024-pitch act
# buffers, four N tiles of 248, the two-launch conv + ChannelNorm plan) and a stand-alone ResidualBlock
# ----------------------------------------------------------------------------------------------------------------------
def test_generator_with_sample_noise_and_standalone_residual_block():
    from hific_b200.network import generator
    from oracle import hific_oracle as O

    def rel(a, b):
        return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
    torch.manual_seed(21)
    gen = generator.Generator((220, 8, 8), 2, C=220, n_residual_blocks=2, sample_noise=True, noise_dim=32)
    with torch.no_grad():
        for name, p in gen.named_parameters():
            if "gamma" in name or "beta" in name or "bias" in name:
                p.add_(0.1 * torch.randn(p.shape))
    sdg = {"Generator." + k: v.detach().clone() for k, v in gen.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    y_hat = torch.round(torch.randn((2, 220, 8, 8), generator=g) * 2)
    z = torch.randn((2, 32, 8, 8), generator=g)
    up = torch.randn((2, 3, 128, 128), generator=g)
    sd2 = {k: v.clone().requires_grad_(True) for k, v in sdg.items()}
    yo = y_hat.clone().requires_grad_(True)
    want = O.generator_forward(sd2, yo, n_residual_blocks=2, noise=z)
    (want * up).sum().backward()
    orig = torch.randn
    torch.randn = lambda *a, **k: z.clone()
    try:
        gen.cuda().eval()
        with torch.no_grad():
            got = gen(y_hat.cuda())
        assert rel(got, want.detach()) < 1.5e-3
        gen.train()
        yc = y_hat.cuda().requires_grad_(True)
        out = gen(yc)
        (out * up.cuda()).sum().backward()
    finally:
        torch.randn = orig
    assert rel(out.detach(), want.detach()) < 1.5e-3
    worst = max(rel(p.grad, sd2["Generator." + k].grad) for k, p in gen.named_parameters())
    assert worst < 5e-2 and rel(yc.grad, yo.grad) < 5e-2, worst
    # stand-alone block (generator.py:33-44), no-grad and autograd
    blk = generator.ResidualBlock((2, 128, 8, 8))
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(0.05 * torch.randn(p.shape))
    sdb = {"b." + k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    x = torch.randn((2, 128, 8, 8), generator=g)
    upb = torch.randn((2, 128, 8, 8), generator=g)
    xo = x.clone().requires_grad_(True)
    wb = O.residual_block(sdb, "b", xo)
    (wb * upb).sum().backward()
    blk.cuda().eval()
    with torch.no_grad():
        assert rel(blk(x.cuda()), wb.detach()) < 1.5e-3
    blk.train()
    xc = x.cuda().requires_grad_(True)
    (blk(xc) * upb.cuda()).sum().backward()
    assert max(rel(p.grad, sdb["b." + k].grad) for k, p in blk.named_parameters()) < 3e-2
    assert rel(xc.grad, xo.grad) < 3e-2
