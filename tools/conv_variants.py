"""Times the dominant conv (960->960 3x3 on 16x16 maps, batch 32) under every cluster / pair mode.
L2 is flushed before each launch.  Run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hific_b200.ops import Conv, Geom, OUT_NHWC_F32, PAD_REFLECT

B = int(os.environ.get("HFC_B", 32))
g = Geom(B, 16, 16, 960, 960, 1, 1, 1, 1)
x = (torch.randn(g.shape, device="cuda") * 0.5).half()
w = torch.randn(960, 960, 3, 3, device="cuda") * 0.01
b = torch.randn(960, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for name, kw in [("1x1", dict(cluster=(1, 1))), ("2x1 multicast", dict(cluster=(2, 1), pair=2)),
                 ("1x2 multicast", dict(cluster=(1, 2))), ("2x2 multicast", dict(cluster=(2, 2), pair=2)),
                 ("2x1 pair", dict(cluster=(2, 1), pair=1)), ("2x2 pair+mc", dict(cluster=(2, 2), pair=1)),
                 ("2x1 pair bn192", dict(cluster=(2, 1), pair=1, block_n=192)),
                 ("2x1 pair bn160", dict(cluster=(2, 1), pair=1, block_n=160)),
                 ("2x2 pair bn160", dict(cluster=(2, 2), pair=1, block_n=160)),
                 ("2x1 pair bn256", dict(cluster=(2, 1), pair=1, block_n=256)),
                 ("2x1 pair bn128", dict(cluster=(2, 1), pair=1, block_n=128))]:
    try:
        conv = Conv(g, 960, 3, pad_mode=PAD_REFLECT, pad=(1, 1, 1, 1), out_mode=OUT_NHWC_F32, **kw)
        out = conv.alloc_out("cuda")
        for _ in range(3):
            conv(x, w, b, out=out)
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        err = ((out - ref).norm() / ref.norm()).item()
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            conv(x, w, b, out=out)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = 100.0 * tot
        print(f"{name:18s} stages {conv.info.stages} bn {conv.info.block_n:3d}x{conv.info.n_tiles} nsub {conv.info.nsub}  {us:7.1f} us  "
              f"{conv.flops / us / 1e6:7.1f} TFLOP/s  rel.diff vs 1x1 {err:.1e}")
    except Exception as e:  # keep going: one broken mode must not hide the others
        print(f"{name:18s} FAILED: {e}")
        break
