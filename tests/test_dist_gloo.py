"""world_size-2 test (gloo, CPU) of the multi-GPU plumbing of the training step: batch sharding, the coalesced gradient
all-reduce and the max-over-ranks timing rule.  The CUDA kernels are not involved (they have no CPU path); the
helpers are backend-agnostic and run over NCCL on the GPU box (bench.py)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from hific_b200.dist import allreduce_gradients, max_over_ranks, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(5, 2, 1))


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(11)
        x = torch.randn(8, 3, 6, 6)                       # the global batch, identical on every rank
        lo, hi = shard_range(x.shape[0], rank, world)
        m = _model()
        frozen = list(m.parameters())[-1]
        loss = m(x[lo:hi]).square().mean()
        loss.backward()
        frozen.grad = None                                # a parameter the step did not touch (same on all ranks)
        nbytes = allreduce_gradients(list(m.parameters()))
        slow = max_over_ranks(10.0 + rank, torch.device("cpu"))
        out.put((rank, [None if p.grad is None else p.grad.clone() for p in m.parameters()], nbytes, slow, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (1, 7, 8, 32):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gradient_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: the full-batch gradient on one process (equal shards: mean of the shard means == full mean)
    torch.manual_seed(11)
    x = torch.randn(8, 3, 6, 6)
    m = _model()
    m(x).square().mean().backward()
    ref = [p.grad for p in m.parameters()]
    assert [r[4] for r in results] == [(0, 4), (4, 8)]
    for rank, grads, nbytes, slow, _ in results:
        assert slow == 11.0                                        # max over ranks of (10 + rank)
        assert grads[-1] is None                                   # untouched parameter stays without gradient
        assert nbytes == sum(g.numel() for g in grads[:-1]) * 4    # one flat fp32 buffer
        for g, r in zip(grads[:-1], ref[:-1]):
            assert torch.allclose(g, r, rtol=1e-5, atol=1e-7)


# ----------------------------------------------------------------------------------------------------------------------
# In-backward reducer on the REAL training plans (kernels emulated on the CPU): the gradients of every layer are handed to
# the reducer from inside the network Functions, bucketed, all-reduced while the backward is still walking
# ----------------------------------------------------------------------------------------------------------------------
def _plan_worker(rank, world, port, out):
    import logging
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from emulation import train_step_cpu_emulation
    from hific_b200 import synth
    from hific_b200.config import mse_lpips_args
    from hific_b200.dist import InBackwardGradientReducer
    from hific_b200.model import Model
    from oracle.ref_shim import NoiseFeeder
    if world > 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg = mse_lpips_args()
        cfg.n_residual_blocks = 1
        m = Model(cfg, logging.getLogger(f"rank{rank}"))
        m.load_state_dict(synth.synth_state_dict(0, n_residual_blocks=1), strict=True)
        m.train()
        n_total = 4
        x = synth.synth_image(n_total, 128, 128, 0)
        nz, ny = synth.synth_noise((n_total, 320, 2, 2), "zd", 0), synth.synth_noise((n_total, 220, 8, 8), "yd", 0)
        lo, hi = shard_range(n_total, rank, world)
        reducer = InBackwardGradientReducer(dist, world, bucket_bytes=4 << 20)
        with train_step_cpu_emulation():
            for step in range(2):     # step 0 calibrates the loss scales (its gradients are handed over per Function),
                for p in m.parameters():      # step 1 is the steady state: layer by layer from inside the Functions
                    p.grad = None
                with NoiseFeeder([nz[lo:hi], ny[lo:hi]]):
                    inter, info = m.compression_forward(x[lo:hi])
                loss = 2.0 * inter.n_bpp + cfg.k_M * m.distortion_loss(inter.reconstruction, inter.input_image)
                with reducer:
                    loss.backward()
        density = list(m.Hyperprior.hyperlatent_likelihood.parameters())
        reducer.reduce_rest(density)
        grads = {k: p.grad.numpy().copy() for k, p in m.named_parameters() if p.grad is not None}   # plain arrays: no fd passing
        emitted = sum(p.numel() * 4 for k, p in m.named_parameters() if "hyperlatent_likelihood" not in k)
        out.put((rank, grads, reducer.bytes_reduced, reducer.buckets_launched, emitted))
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_in_backward_reducer_on_training_plans_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, out)) for r in range(2)]
    procs.append(ctx.Process(target=_plan_worker, args=(0, 1, 0, out)))       # the full batch on one process
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in range(3)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = [r for r in res if r[3] == 0]
    multi = sorted([r for r in res if r[3] > 0], key=lambda t: t[0])
    assert len(single) == 1 and len(multi) == 2
    ref = single[0][1]
    for rank, grads, nbytes, buckets, emitted in multi:
        assert nbytes == emitted                          # every plan parameter went through the in-backward path, once
        assert buckets >= 8                               # several buckets per network, not one per network
        assert set(grads) == set(ref)
        for k in ref:
            a, b = torch.from_numpy(grads[k]).double(), torch.from_numpy(ref[k]).double()
            # identical on both ranks; equal to the single-process full-batch gradient up to the fp16 operand rounding
            # (each rank scales / rounds its own half) -- the density parameters went through reduce_rest
            assert (grads[k] == multi[0][1][k]).all(), k
            assert float((a - b).norm() / b.norm().clamp_min(1e-30)) < 2e-2, k


def _then_step_worker(rank, world, port, out):
    import torch.distributed as dist
    from hific_b200.dist import allreduce_then_step
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(11)
        x = torch.randn(8, 3, 6, 6)
        lo, hi = shard_range(x.shape[0], rank, world)
        m = _model()
        opt = torch.optim.Adam(m.parameters(), lr=1e-2)   # no step_subset: the helper must take the plain path
        m(x[lo:hi]).square().mean().backward()
        buckets = allreduce_then_step(opt, dist, world)
        out.put((rank, buckets, [p.detach().numpy().copy() for p in m.parameters()]))
    finally:
        dist.destroy_process_group()


def test_allreduce_then_step_falls_back_to_allreduce_plus_step():
    """hific_b200.dist.allreduce_then_step with an optimizer that cannot step bucket by bucket (and on gloo): exactly
    allreduce_gradients + optimizer.step() -- the parameters of both ranks equal one process stepping on the mean gradient
    of the two shards.  (The bucketed NCCL path is self-checked by bench.py on the GPUs.)"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_then_step_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    torch.manual_seed(11)
    x = torch.randn(8, 3, 6, 6)
    m = _model()
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    grads = []
    for r in range(world):
        lo, hi = shard_range(8, r, world)
        m.zero_grad()
        m(x[lo:hi]).square().mean().backward()
        grads.append([p.grad.clone() for p in m.parameters()])
    for p, g0, g1 in zip(m.parameters(), *grads):
        p.grad = (g0 + g1) / world
    opt.step()
    for rank, buckets, params in results:
        assert buckets == 0
        for got, want in zip(params, m.parameters()):
            assert torch.allclose(torch.from_numpy(got), want.detach(), rtol=1e-5, atol=1e-7)
