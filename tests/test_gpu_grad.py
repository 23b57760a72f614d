"""GPU tests of the backward building blocks against torch autograd (fp32, TF32 off) on the same device.
Operands are pre-rounded (activations / weights to fp16, upstream gradients to bf16) so that the remaining
difference is accumulation order only."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from hific_b200 import ops  # noqa: E402
from hific_b200.grad import ConvGrad, gemm_nt  # noqa: E402
from hific_b200.ops import Geom, PAD_REFLECT, PAD_ZERO, round_up  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_gemm_nt_rejects_mixed_formats():
    """tcgen05 kind::f16 raises an illegal-instruction fault for fp16 x bf16 operand pairs (measured on B200), so the
    library refuses them up front instead of poisoning the CUDA context."""
    a = torch.zeros(128, 64, dtype=torch.int16, device=DEV)
    with pytest.raises(RuntimeError, match="mixed"):
        gemm_nt(a, True, a, False, 128, 128, 64)


@pytest.mark.parametrize("fmt", ["f16", "bf16"])
@pytest.mark.parametrize("m,n,k,splits", [(128, 64, 256, 0), (960, 8640 // 8, 2048, 0), (60, 392, 64 * 700, 0),
                                          (200, 100, 64 * 37, 5)])
def test_gemm_nt(m, n, k, splits, fmt):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(m, k, generator=g).to(DEV)
    b = torch.randn(n, k, generator=g).to(DEV)
    a_bf, b_bf = fmt in ("bf16", "mixed"), fmt == "bf16"
    a16 = a.bfloat16() if a_bf else a.half()
    b16 = b.bfloat16() if b_bf else b.half()
    ref = a16.float() @ b16.float().t()
    out = gemm_nt(a16.view(torch.int16), a_bf, b16.view(torch.int16), b_bf, m, n, k, k_splits=splits)
    torch.cuda.synchronize()
    assert rel(out[:, :n], ref) < 2e-5


def run_grad_case(n, cin, h, w, cout, k, stride=1, pad=(0, 0, 0, 0), pad_mode=PAD_ZERO, transposed=False, window=False,
                  seed=0):
    g = torch.Generator().manual_seed(seed)
    # values representable in both fp16 and bf16 so that the operand conversions are exact
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float().to(DEV).requires_grad_(True)
    if transposed:
        wt = (torch.randn(cin, cout, k, k, generator=g) / math.sqrt(cin * k * k)).bfloat16().float().to(DEV)
    else:
        wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).bfloat16().float().to(DEV)
    wt.requires_grad_(True)
    bias = torch.zeros(cout, device=DEV, requires_grad=True)
    pt, pl, pb, pr = pad
    if transposed:
        y = F.conv_transpose2d(x, wt, bias, stride=stride, padding=pt, output_padding=stride - 1)
    else:
        xp = F.pad(x, (pl, pr, pt, pb), mode="reflect" if pad_mode == PAD_REFLECT else "constant")
        y = F.conv2d(xp, wt, bias, stride=stride)
    dy = torch.randn(y.shape, generator=g).bfloat16().float().to(DEV)
    y.backward(dy)

    cpad = 8 if window else round_up(cin, 64)
    border = (pt, pl, pb, pr + (1 if window else 0)) if pad_mode == PAD_REFLECT else (0, 0, 0, 0)
    in_geom = Geom(n, h, w, cin, cpad, *border)
    x_act = ops.nchw_to_act(x.detach(), in_geom, reflect=(pad_mode == PAD_REFLECT))
    cg = ConvGrad(in_geom, cout, k, stride=stride, transposed=transposed, pad_mode=pad_mode, pad=pad)
    dy_rows = dy.permute(0, 2, 3, 1).reshape(-1, cout).contiguous()
    if cout % 4:
        dy_rows = F.pad(dy_rows, (0, 4 - cout % 4)).contiguous()
    dw = cg.weight_grad(x_act, dy_rows)
    db = cg.bias_grad(dy_rows)
    torch.cuda.synchronize()
    assert rel(dw, wt.grad) < 1e-4, f"wgrad rel err {rel(dw, wt.grad)}"
    assert rel(db, bias.grad) < 1e-5
    if not window:
        dx = cg.data_grad(dy_rows, wt.detach())
        torch.cuda.synchronize()
        dx_nchw = dx.view(n, h, w, -1)[..., :cin].permute(0, 3, 1, 2)
        assert rel(dx_nchw, x.grad) < 1e-4, f"dgrad rel err {rel(dx_nchw, x.grad)}"


def test_grad_conv_s1_reflect():
    run_grad_case(2, 64, 16, 16, 96, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT)


def test_grad_conv_s1_reflect_resblock_shape():
    run_grad_case(2, 960, 16, 16, 960, 3, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT)


def test_grad_conv_s1_zero_pad_220_320():
    run_grad_case(2, 220, 16, 16, 320, 3, pad=(1, 1, 1, 1), pad_mode=PAD_ZERO)


def test_grad_conv_s2_asym_reflect():
    run_grad_case(2, 60, 32, 32, 120, 3, stride=2, pad=(1, 0, 0, 1), pad_mode=PAD_REFLECT)


def test_grad_conv_5x5_s2_reflect():
    run_grad_case(3, 320, 16, 16, 320, 5, stride=2, pad=(2, 2, 2, 2), pad_mode=PAD_REFLECT)


def test_grad_convT_k3_s2():
    run_grad_case(2, 240, 16, 16, 120, 3, stride=2, pad=(1, 1, 1, 1), transposed=True)


def test_grad_convT_k5_s2_and_k3_s1():
    run_grad_case(3, 320, 4, 4, 320, 5, stride=2, pad=(2, 2, 2, 2), transposed=True)
    run_grad_case(2, 320, 16, 16, 220, 3, stride=1, pad=(1, 1, 1, 1), transposed=True)


def test_grad_first_layer_window_and_tiny_cout_head():
    run_grad_case(2, 3, 32, 32, 60, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT, window=True)   # wgrad only (no dx needed)
    run_grad_case(2, 60, 24, 40, 3, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT)               # swap formulation
    run_grad_case(1, 60, 10, 136, 3, 7, pad=(3, 3, 3, 3), pad_mode=PAD_REFLECT)              # wide: flat-segment dgrad


def test_grad_conv_4x4_s2_discriminator():
    run_grad_case(2, 64, 32, 32, 128, 4, stride=2, pad=(1, 1, 1, 1), pad_mode=PAD_REFLECT)
