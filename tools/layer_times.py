"""Per-layer device timings (CUDA events, L2 flushed before every launch) of the forward plans at the
bench size.  Prints a table and writes gpurun_out/layer_times.json.  Run on the GPU box."""
import json
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

os.environ.setdefault("HFC_LPIPS_SYNTHETIC", "1")   # no checkpoints on the boxes: seeded stand-in, as the tests do

from hific_b200 import ops, synth
from hific_b200.config import mse_lpips_args
from hific_b200.model import Model
from hific_b200.ops import ACT_NONE, ACT_RELU

B = int(os.environ.get("HFC_B", 32))
REPS = 5
dev = "cuda"
model = Model(mse_lpips_args(), logging.getLogger("lt"))
model.load_state_dict(synth.synth_state_dict(0), strict=True)
model.cuda().eval()
x = synth.synth_image(B, 256, 256, 1).cuda()
with torch.no_grad():
    inter, info = model.compression_forward(x)      # builds all plans, fills buffers
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []


def timeit(name, fn, flops=0.0, bytes_=0.0):
    fn()
    tot = 0.0
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = 1e3 * tot / REPS
    rows.append(dict(name=name, us=us, gflop=flops / 1e9, tflops=flops / us / 1e6 if flops else 0.0,
                     mbytes=bytes_ / 1e6, gbs=bytes_ / us / 1e3 if bytes_ else 0.0))
    print(f"{name:34s} {us:9.1f} us  {flops / 1e9:9.2f} GFLOP {flops / us / 1e6 if flops else 0:8.1f} TFLOP/s"
          f"  {bytes_ / 1e6:8.1f} MB {bytes_ / us / 1e3 if bytes_ else 0:8.1f} GB/s")


def nbytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts))


with torch.no_grad():
    E, G, H = model.Encoder, model.Generator, model.Hyperprior
    ep = E._plans.get(x)
    timeit("E.nchw_to_act", lambda: ops.nchw_to_act(x, ep.g_in, reflect=True, out=ep.x_act), 0, nbytes(x, ep.x_act))
    h = ep.x_act
    for i in range(5):
        blk = getattr(E, f"conv_block{i + 1}")
        conv = ep.convs[i]
        if i in ep.cn:
            timeit(f"E{i + 1}.conv({conv.info.block_n}x{conv.info.n_tiles})", lambda: conv(h, blk[1].weight, blk[1].bias, out=ep.bufs[i]),
                   conv.flops, nbytes(h, ep.bufs[i]))
            timeit(f"E{i + 1}.channelnorm", lambda: ops.channelnorm(ep.bufs[i], ep.cn[i], blk[2].gamma, blk[2].beta, act=ACT_RELU,
                                                                   reflect=True, out_act=ep.cn_bufs[i]), 0, nbytes(ep.bufs[i], ep.cn_bufs[i]))
            h = ep.cn_bufs[i]
        else:
            timeit(f"E{i + 1}.conv+cn({conv.info.block_n})", lambda: conv(h, blk[1].weight, blk[1].bias, blk[2].gamma, blk[2].beta, out=ep.bufs[i]),
                   conv.flops, nbytes(h, ep.bufs[i]))
            h = ep.bufs[i]
    last = E.conv_block_out[1]
    timeit("E6.conv(nchw out)", lambda: ep.convs[5](h, last.weight, last.bias), ep.convs[5].flops)
    y = E(x)
    hp = H.analysis_net._plans.get(y)
    timeit("H.analysis (3 convs + to_act)", lambda: H.analysis_net(y), hp.flops)
    z = H.analysis_net(y)
    timeit("H.synthesis_mu (9 launches)", lambda: H.synthesis_mu(z), H.synthesis_mu._plans.get(z).flops)
    timeit("H.forward (all)", lambda: H(y, spatial_shape=(256, 256)))
    mu, sg = H.synthesis_mu(z), H.synthesis_std(z)
    ny = torch.rand_like(y) - 0.5
    timeit("H.latent_likelihood", lambda: ops.latent_likelihood(y, mu, sg, ny), 0, nbytes(y) * 5)
    timeit("H.hyperlatent_likelihood", lambda: ops.hyperlatent_likelihood(z, H.hyperlatent_likelihood.packed_params(), torch.rand_like(z) - .5), 0, nbytes(z) * 4)
    yh = info.decoded
    gp = G._plans.get(yh)
    init = G.conv_block_init
    timeit("G.nchw_to_act+cn", lambda: ops.nchw_to_act(yh, gp.g_in, reflect=True, norm=True, gamma=init[0].gamma, beta=init[0].beta, out=gp.in_act))
    timeit("G0.conv", lambda: gp.conv_init(gp.in_act, init[2].weight, init[2].bias, out=gp.rows), gp.conv_init.flops)
    blk = G.resblock_0
    c1 = gp.res_convs[0][0]
    timeit("G.res conv 960->960 (240x4)", lambda: c1(gp.act_a, blk.conv1.weight, blk.conv1.bias, out=gp.rows), c1.flops)
    timeit("G.res channelnorm(+res)", lambda: ops.channelnorm(gp.rows, gp.g_b1, blk.norm2.gamma, blk.norm2.beta, act=ACT_NONE, reflect=True,
                                                             res1=gp.head_f32, want_f32=True, out_f32=gp.x_f32[0], out_act=gp.act_a),
           0, nbytes(gp.rows, gp.head_f32, gp.x_f32[0], gp.act_a))
    hcur = gp.act_flat
    for i, (conv, g_cn, out_buf, cn_buf) in enumerate(gp.ups):
        ub = getattr(G, f"upconv_block{i + 1}")
        if g_cn is None:
            timeit(f"G.up{i + 1} convT+cn ({conv.info.block_n}) 4 phases", lambda: conv(hcur, ub[0].weight, ub[0].bias, ub[1].gamma, ub[1].beta, out=out_buf),
                   conv.flops, nbytes(hcur, out_buf))
            hcur = out_buf
        else:
            timeit(f"G.up{i + 1} convT ({conv.info.block_n}x{conv.info.n_tiles}) 4 phases", lambda: conv(hcur, ub[0].weight, ub[0].bias, out=out_buf), conv.flops,
                   nbytes(hcur, out_buf))
            timeit(f"G.up{i + 1} channelnorm", lambda: ops.channelnorm(out_buf, g_cn, ub[1].gamma, ub[1].beta, act=ACT_RELU, out_act=cn_buf), 0, nbytes(out_buf, cn_buf))
            hcur = cn_buf
    lastg = G.conv_block_out[1]
    timeit("G3.conv 7x7 60->3", lambda: gp.conv_out(hcur, lastg.weight, lastg.bias), gp.conv_out.flops, nbytes(hcur))
    timeit("Encoder total", lambda: E(x), ep.flops)
    timeit("Generator total", lambda: G(yh), gp.flops)
    timeit("Model.compression_forward", lambda: model.compression_forward(x), B * 99.89e9)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/layer_times.json", "w"), indent=1)
