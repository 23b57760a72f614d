"""Which submission of PipelinedForward differs from the synchronous forward, and does it equal another submission's result
(slot mix-up) or nothing at all (a race inside the forward)?  Repeats the comparison 20 times in one process."""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HFC_LPIPS_SYNTHETIC", "1")
import torch

from hific_b200 import synth
from hific_b200.config import ModelModes, mse_lpips_args
from hific_b200.model import Model
from hific_b200.pipeline import PipelinedForward

m = Model(mse_lpips_args(), logging.getLogger("pipe"), model_mode=ModelModes.EVALUATION)
m.load_state_dict(synth.synth_state_dict(0), strict=False)
m.cuda().eval()
xs = [synth.synth_image(2, 128, 128, 20 + i).pin_memory() for i in range(5)]
bad = 0
for rep in range(20):
    with torch.no_grad():
        ref = [m(x.cuda(), writeout=False)[0].cpu().clone() for x in xs]
        ref2 = [m(x.cuda(), writeout=False)[0].cpu().clone() for x in xs]
    sync_ok = all(torch.equal(a, b) for a, b in zip(ref, ref2))
    pipe = PipelinedForward(m, depth=2)
    tickets, got = [], []
    for i, x in enumerate(xs):
        tickets.append(pipe.submit(x))
        if i >= 1:
            got.append(pipe.result(tickets[i - 1])[0].clone())
    got.append(pipe.result(tickets[-1])[0].clone())
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(ref, got)):
        if not torch.equal(a, b):
            bad += 1
            same_as = [j for j, r in enumerate(ref) if torch.equal(r, b)]
            frac = float((a != b).float().mean())
            print(f"rep {rep}: submission {i} differs (fraction of elements {frac:.4f}, max abs {float((a - b).abs().max()):.4f}); "
                  f"equals the synchronous result of inputs {same_as}; synchronous run-to-run equal: {sync_ok}")
print("mismatching submissions:", bad, "of", 20 * len(xs))
