"""Launches the compress-path kernels (csrc/symbols.cu) and both schedules of the latent-likelihood kernel a few times
at the c5 size (y = 8 x 220 x 64 x 64) and the c2 size so that `ncu --set full -k regex:"symbols|likelihood"` can
capture them.  GPU box only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hific_b200 import ops
from hific_b200._lib import SYM_BATCH_STEPS, SYM_PIXEL_STEPS
from hific_b200.compression.prior_model import prior_scale_table

table = torch.clamp(prior_scale_table(), 0.11).cuda()
for shape, layout in (((8, 220, 64, 64), SYM_BATCH_STEPS), ((1, 220, 64, 64), SYM_PIXEL_STEPS)):
    y = torch.randn(shape, device="cuda") * 3
    mu = torch.randn(shape, device="cuda")
    sc = torch.rand(shape, device="cuda") * 3
    for _ in range(2):
        q = ops.quantize_symbols(y, mu, sc, table, 0.11, "gaussian", layout, want_bits=True)
        ops.scale_indices(sc, table, 0.11, layout)
        ops.dequantize_symbols(q["symbols"], mu, shape, layout)
torch.cuda.synchronize()
n = 1802240
y = torch.randn(n, device="cuda").view(1, 1, 1, n) * 2
mu, s, nz = torch.randn_like(y), torch.rand_like(y) * 2, torch.rand_like(y) - 0.5
for v in ("1", "2", "3", "1", "2", "3"):
    os.environ["HFC_LIKELIHOOD_V"] = v
    ops.latent_likelihood(y, mu, s, nz)
torch.cuda.synchronize()
