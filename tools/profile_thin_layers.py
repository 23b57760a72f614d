"""Launches the three epilogue-bound big-map layers of the forward pass at batch 32 -- E1 (7x7, 3 -> 60 + ChannelNorm + ReLU,
window packing), G.up4 (transposed 3x3 stride 2, 120 -> 60 + ChannelNorm + ReLU, 4 phases) and G3 (7x7, 60 -> 3, tap-in-N)
-- a few times each, so that

    ncu --set full --import-source on -k regex:conv_igemm -c 24 -o gpurun_out/thin python tools/profile_thin_layers.py
    HFC_THIN_EPILOGUE=1 ncu ... -o gpurun_out/thin_on python tools/profile_thin_layers.py

capture them with and without the thin epilogue; also prints CUDA-event times (L2 flushed).  GPU box only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hific_b200.engine import EncoderPlan, GeneratorPlan

B = int(os.environ.get("HFC_B", 32))
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(name, fn, reps=int(os.environ.get('HFC_REPS', 5))):
    for _ in range(min(2, reps)):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name}: {sorted(ts)[len(ts) // 2]:.1f} us (HFC_THIN_EPILOGUE={os.environ.get('HFC_THIN_EPILOGUE', '1 (default)')})")


THIN = os.environ.get("HFC_THIN_EPILOGUE", "1 (default)")
enc = EncoderPlan(B, 256, 256, 3, 220, dev)
x_act = (torch.rand(enc.g_in.shape, device=dev) - 0.5).half()
w1 = torch.randn(60, 3, 7, 7, device=dev) * 0.05
b1, g1, be1 = torch.randn(60, device=dev), torch.rand(60, device=dev) + 0.5, torch.randn(60, device=dev)
timed("E1   7x7 3->60 +CN+ReLU", lambda: enc.convs[0](x_act, w1, b1, g1, be1, out=enc.bufs[0]))

gen = GeneratorPlan(B, 16, 16, 220, 9, 3, dev)
conv4, _, out4, _ = gen.ups[3]
a4 = (torch.randn(conv4.in_geom.shape, device=dev) * 0.5).half()
w4 = torch.randn(120, 60, 3, 3, device=dev) * 0.03
b4, g4, be4 = torch.randn(60, device=dev), torch.rand(60, device=dev) + 0.5, torch.randn(60, device=dev)
timed("G.up4 convT 120->60 +CN+ReLU (4 phases)", lambda: conv4(a4, w4, b4, g4, be4, out=out4))

a5 = (torch.randn(gen.conv_out.in_geom.shape, device=dev) * 0.5).half()
w5 = torch.randn(3, 60, 7, 7, device=dev) * 0.02
b5 = torch.randn(3, device=dev)
timed("G3   7x7 60->3 (tap-in-N)", lambda: gen.conv_out(a5, w5, b5))

# the two N = 128 layers of the same kind (one N tile, big maps): thin only with HFC_THIN_EPILOGUE=2
a1 = (torch.randn(enc.convs[1].in_geom.shape, device=dev) * 0.5).half()
w2 = torch.randn(120, 60, 3, 3, device=dev) * 0.04
b2, g2, be2 = torch.randn(120, device=dev), torch.rand(120, device=dev) + 0.5, torch.randn(120, device=dev)
timed("E2   3x3 s2 60->120 +CN+ReLU", lambda: enc.convs[1](a1, w2, b2, g2, be2, out=enc.bufs[1]))
conv3, _, out3, _ = gen.ups[2]
a3 = (torch.randn(conv3.in_geom.shape, device=dev) * 0.5).half()
w3 = torch.randn(240, 120, 3, 3, device=dev) * 0.02
b3, g3, be3 = torch.randn(120, device=dev), torch.rand(120, device=dev) + 0.5, torch.randn(120, device=dev)
timed("G.up3 convT 240->120 +CN+ReLU (4 phases)", lambda: conv3(a3, w3, b3, g3, be3, out=out3))
