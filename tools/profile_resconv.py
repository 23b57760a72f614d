"""Launches the dominant kernel (Generator residual conv 960->960 3x3, batch 32, auto mode = CTA pairs) a few
times so that `ncu -k regex:conv_igemm -s 3 -c 1 --set full` can capture one warm launch.  GPU box only."""
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
conv = Conv(g, 960, 3, pad_mode=PAD_REFLECT, pad=(1, 1, 1, 1), out_mode=OUT_NHWC_F32)
print("mode: cluster", conv.info.cluster_m, conv.info.cluster_n, "pair", conv.info.pair, "stages", conv.info.stages)
out = conv.alloc_out("cuda")
for _ in range(6):
    conv(x, w, b, out=out)
torch.cuda.synchronize()
# likelihood kernel at c2 and c5 sizes
from hific_b200 import ops
for n in (1802240, 7208960):
    y = torch.randn(n, device="cuda").view(1, 1, 1, n) * 2
    mu = torch.randn_like(y)
    s = torch.rand_like(y) * 2
    nz = torch.rand_like(y) - 0.5
    for _ in range(3):
        ops.latent_likelihood(y, mu, s, nz)
torch.cuda.synchronize()
