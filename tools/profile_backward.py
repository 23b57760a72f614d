"""Launches the dominant BACKWARD kernels at the bench size so that ncu can capture one warm launch of each:

    ncu --set full --clock-control none --import-source on -k regex:"wgrad_igemm|channelnorm_bwd|conv_igemm" \
        -s 12 -c 6 -o gpurun_out/bwd python tools/profile_backward.py

  * implicit weight gradient of the Generator residual conv (960 -> 960, 3x3, batch 32: M=960, N=9x960, K=8192 pixels)
  * its data gradient (3x3 conv over the 18x18 padded domain, free-form 18x7 tiles)
  * ChannelNorm backward at 960 channels (16x16 maps) and at 60 channels (256x256 maps)
GPU box only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hific_b200 import ops
from hific_b200.grad import ConvGrad
from hific_b200.ops import ACT_RELU, Geom, PAD_REFLECT
from hific_b200.train_plan import norm_bwd

B = int(os.environ.get("HFC_B", 32))
dev = "cuda"
g = Geom(B, 16, 16, 960, 960, 1, 1, 1, 1)
x = (torch.randn(g.shape, device=dev) * 0.5).half()
w = torch.randn(960, 960, 3, 3, device=dev) * 0.01
cg = ConvGrad(g, 960, 3, pad_mode=PAD_REFLECT, pad=(1, 1, 1, 1))
dy = torch.randn(B * 256, 960, device=dev) * 1e-3
for _ in range(4):
    dy_act = cg.dy_to_act(dy)
    cg.weight_grad(x, None, dy_act=dy_act)
    cg.data_grad(None, w, dy_act=dy_act)
torch.cuda.synchronize()
for c, hw in ((960, 16), (60, 256)):
    npix = B * hw * hw
    z = torch.randn(npix, c, device=dev)
    gr = torch.randn(npix, c, device=dev)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    for _ in range(4):
        norm_bwd(z, gr, gamma, beta, ACT_RELU)
torch.cuda.synchronize()
print("launches", ops.launch_count())
