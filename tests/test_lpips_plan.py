"""CPU: the native LPIPS trunk PLAN (hific_b200.loss.lpips_trunk: geometry, space-to-depth re-indexing of the 11x11
stride-4 head, batching of target and reconstruction, order of the adjoints) against torchvision's AlexNet + the
oracle's LPIPS + torch autograd, with the CUDA entry points replaced by independent torch stand-ins
(tests/emulation.py).  Tolerance: the plan stores features in fp16 and multiplies fp16 (forward) / bf16 (backward)
operands, so distances agree to 2e-3 relative and input gradients to 5e-2 relative L2 (bf16 gradients, ReLU / arg-max flips)."""
import pytest
import torch

from emulation import lpips_cpu_emulation
from hific_b200.loss import lpips_trunk
from hific_b200.loss.perceptual import PerceptualLoss
from oracle import hific_oracle as O


def test_s2d_weights_reproduce_the_stride4_head():
    g = torch.Generator().manual_seed(0)
    w = torch.randn((64, 3, 11, 11), generator=g)
    x = torch.randn((2, 3, 70, 83), generator=g)
    ref = torch.nn.functional.conv2d(x, w, stride=4, padding=2)
    hs, ws = ref.shape[2] + 2, ref.shape[3] + 2
    xp = torch.zeros(2, 3, 4 * hs, 4 * ws)
    xp[:, :, 2:72, 2:85] = x
    s2d = xp.view(2, 3, hs, 4, ws, 4).permute(0, 3, 5, 1, 2, 4).reshape(2, 48, hs, ws)
    got = torch.nn.functional.conv2d(s2d, lpips_trunk.s2d_weights(w))
    assert got.shape == ref.shape and (got - ref).abs().max() < 1e-3


@pytest.mark.parametrize("n,h,w,normalize", [(2, 128, 128, True), (1, 100, 144, False)])
def test_native_plan_matches_torchvision_and_autograd(n, h, w, normalize):
    loss = PerceptualLoss()
    g = torch.Generator().manual_seed(n + h)
    target = torch.rand((n, 3, h, w), generator=g)
    pred = (target + 0.1 * torch.randn((n, 3, h, w), generator=g)).clamp(0, 1)
    pred_ref = pred.clone().requires_grad_(True)
    want = O.lpips_forward(loss.trunk, [p.detach() for p in loss.lins], pred_ref, target, normalize=normalize).view(-1)
    up = torch.linspace(0.5, 1.5, n)
    (want * up).sum().backward()
    plan = lpips_trunk.LpipsTrunkPlan(n, h, w, torch.device("cpu"))
    pred_nat = pred.clone().requires_grad_(True)
    with lpips_cpu_emulation():
        got = lpips_trunk.LpipsTrunkFn.apply(pred_nat, target, plan, loss, normalize)
        (got * up).sum().backward()
    assert torch.allclose(got, want.detach(), rtol=2e-3, atol=1e-6), (got, want)
    gw, gg = pred_ref.grad, pred_nat.grad
    rel = ((gg - gw).norm() / gw.norm()).item()
    assert rel < 5e-2, rel          # same bar as the other networks (tests/test_gpu_train.py): ReLU / arg-max flips of fp16 features
    assert plan.flops > 0
