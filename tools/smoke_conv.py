"""Smallest possible GPU exercise of the tcgen05 conv kernel, with verbose diagnostics.
Run on the GPU box before the test-suite so that a descriptor/layout bug shows up as numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from hific_b200 import ops
from hific_b200.ops import Conv, Geom, PAD_REFLECT, PAD_ZERO, OUT_NCHW_F32

torch.backends.cudnn.allow_tf32 = False
print("device", torch.cuda.get_device_name(0), ops.device_info())
torch.manual_seed(0)
for (n, cin, h, w, cout, k, pad, mode) in [(1, 64, 16, 8, 16, 1, 0, PAD_ZERO), (2, 64, 16, 16, 64, 3, 1, PAD_REFLECT),
                                           (2, 128, 16, 16, 240, 3, 1, PAD_ZERO)]:
    x = torch.randn(n, cin, h, w).half().float().cuda()
    wt = (torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5).half().float().cuda()
    b = torch.randn(cout).cuda()
    g = Geom(n, h, w, cin, 64 * ((cin + 63) // 64), *( (pad,) * 4 if mode == PAD_REFLECT else (0,) * 4))
    xa = ops.nchw_to_act(x, g, reflect=(mode == PAD_REFLECT))
    torch.cuda.synchronize()
    chk = g.interior(xa)
    print("to_act interior max err", (chk - x).abs().max().item())
    conv = Conv(g, cout, k, pad_mode=mode, pad=(pad,) * 4, out_mode=OUT_NCHW_F32)
    print("info: block_n", conv.info.block_n, "n_tiles", conv.info.n_tiles, "m_tiles", conv.info.m_tiles,
          "stages", conv.info.stages, "k_total", conv.info.k_total)
    out = conv(xa, wt, b)
    torch.cuda.synchronize()
    xp = F.pad(x, (pad,) * 4, mode="reflect" if mode == PAD_REFLECT else "constant")
    ref = F.conv2d(xp, wt, b)
    err = (out - ref).abs().max().item()
    print(f"conv n={n} cin={cin} {h}x{w} cout={cout} k={k}: max abs err {err:.3e}  ref absmax {ref.abs().max().item():.3f}")
    if err > 1e-3:
        print("out[0,0,:2,:8]", out[0, 0, :2, :8].tolist())
        print("ref[0,0,:2,:8]", ref[0, 0, :2, :8].tolist())
print("SMOKE DONE")
