"""Multi-GPU plumbing of the training step (one process per GPU, torch.distributed): the HiFIC path is purely
data-parallel (ChannelNorm is per pixel, no batch statistics), so ranks only exchange GRADIENTS -- one coalesced
all-reduce of the stepped parameter group after backward (`allreduce_gradients`), or bucketed all-reduces issued from
inside the backward (`InBackwardGradientReducer`; train.py has no DDP wrapper; north_star asks for "NCCL allreduce over
NVLink for gradients only").  Backend-agnostic: NCCL on the GPUs, gloo in the CPU tests."""
import os

import torch


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) slice of n_items owned by `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params, dist=None, world=None):
    """Average `.grad` of `params` over all ranks with ONE collective launch.  Parameters without a
    gradient on this rank are skipped on EVERY rank only if they have none anywhere -- callers pass the parameter group
    that the step back-propagated into, which is the same on all ranks.  Returns the number of bytes reduced."""
    if dist is None:
        import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return 0
    nbytes = sum(g.numel() * g.element_size() for g in grads)
    if dist.get_backend() == "nccl":
        # NCCL: one grouped launch over the gradient tensors IN PLACE (ncclGroupStart/End via torch's coalescing
        # manager) with ncclAvg -- no flatten / unflatten copies (3 extra passes over 726 MB) and no division kernel
        with dist._coalescing_manager(device=grads[0].device):
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.AVG)
        return nbytes
    flat = torch._utils._flatten_dense_tensors(grads)          # gloo (CPU tests): no AVG, no coalescing fast path
    dist.all_reduce(flat)
    flat.div_(world)
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)
    return nbytes


_step_comm = {}


def allreduce_then_step(optimizer, dist=None, world=None, bucket_bytes=None):
    """The plain after-backward gradient all-reduce PIPELINED with the optimizer: the stepped group's gradients are cut
    into buckets (HFC_STEP_BUCKET_MB, default 64 MB), every bucket is one coalesced in-place ncclAvg on a communication
    stream, and the Adam launch of bucket i (`hific_b200.optim.Adam.step_subset`, compute stream) waits for bucket i only --
    so the 0.8 ms HBM-bound Adam pass runs underneath the NVLink-bound collective of the next buckets instead of after
    all of it.  Unlike hiding the collective behind the BACKWARD (InBackwardGradientReducer), nothing here competes for
    whole SMs: Adam is an elementwise grid that co-resides with NCCL's CTAs.  Same arithmetic per parameter as
    `allreduce_gradients(params); optimizer.step()`.  Optimizers without `step_subset`, non-NCCL backends and world 1
    take exactly that plain path."""
    if dist is None:
        import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    groups = [(g, [p for p in g["params"] if p.grad is not None]) for g in optimizer.param_groups]
    params = [p for _, ps in groups for p in ps]
    if (world == 1 or not params or not hasattr(optimizer, "step_subset") or dist.get_backend() != "nccl" or
            not params[0].is_cuda):
        allreduce_gradients(params, dist, world)
        optimizer.step()
        return 0
    if bucket_bytes is None:
        bucket_bytes = int(os.environ.get("HFC_STEP_BUCKET_MB", "64")) << 20
    dev = params[0].device
    comm = _step_comm.get(dev)
    if comm is None:
        comm = _step_comm[dev] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    comm.wait_stream(main)                                   # the gradients are final
    plan = []                                                # (group, params, event)
    for group, ps in groups:
        bucket, nbytes = [], 0
        for i, p in enumerate(ps):
            bucket.append(p)
            nbytes += p.grad.numel() * p.grad.element_size()
            if nbytes >= bucket_bytes or i == len(ps) - 1:
                with torch.cuda.stream(comm):
                    with dist._coalescing_manager(device=dev):
                        for q in bucket:
                            dist.all_reduce(q.grad, op=dist.ReduceOp.AVG)
                    ev = torch.cuda.Event()
                    ev.record(comm)
                for q in bucket:
                    q.grad.record_stream(comm)
                plan.append((group, bucket, ev))
                bucket, nbytes = [], 0
    for group, bucket, ev in plan:
        main.wait_event(ev)
        optimizer.step_subset(group, bucket)
    return len(plan)


def max_over_ranks(value, device, dist=None):
    """The contract's timing rule: a step takes as long as its slowest rank."""
    if dist is None:
        import torch.distributed as dist
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


class InBackwardGradientReducer:
    """Gradient all-reduce issued from INSIDE the backward of the network Functions.

    Each network of the hot path is ONE autograd Function, so grad-ready hooks (round 1's reducer) only fire
    when a whole network is done -- and the Generator, 87 % of the parameters, is done first, leaving almost nothing to
    hide its 630 MB behind.  The training plans therefore hand their parameter gradients over layer by layer
    (`grad.emit`, called the moment a layer's weight / bias / norm gradients are final and un-scaled): this reducer
    collects them into buckets of `bucket_bytes` and all-reduces (mean) every full bucket IN PLACE on a communication
    stream while the compute stream walks the remaining layers.  At the end of each Function's backward
    (`function_done`) the compute stream waits for the buckets still in flight, because autograd's AccumulateGrad reads
    the returned tensors next; what stays exposed per Function is the tail of its last bucket only.

        reducer = InBackwardGradientReducer(dist, world)
        with reducer:                       # installs itself as grad._grad_sink
            loss.backward()
        reducer.reduce_rest(other_params)   # parameters outside the plans (hyper-latent density, discriminator)

    Every rank walks the same plans in the same order, so buckets line up across ranks without negotiation.
    Backend-agnostic: NCCL (coalesced in-place ncclAvg on a side stream) or gloo (CPU tests: flatten, sum, divide)."""

    def __init__(self, dist=None, world=None, bucket_bytes=None, group=None):
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        if bucket_bytes is None:
            bucket_bytes = int(os.environ.get("HFC_REDUCER_BUCKET_MB", "32")) << 20
        self.bucket_bytes = int(bucket_bytes)
        self.group = group                  # None: the default process group; see reducer_group()
        self.pending, self.pending_bytes = [], 0
        self.inflight = []                  # gloo: (work, flat, grads)
        self._comm, self._done = None, None
        self.bytes_reduced, self.buckets_launched, self.ids = 0, 0, set()
        self.debug_sync = os.environ.get("HFC_REDUCER_SYNC") == "1"      # diagnostics: host-synchronise after every bucket

    # -- sink protocol (hific_b200.grad) ---------------------------------------------------------------------------
    def submit(self, tensors):
        if self.world == 1:
            return
        for t in tensors:
            self.pending.append(t)
            self.pending_bytes += t.numel() * t.element_size()
            self.ids.add(t.data_ptr())
        if self.pending_bytes >= self.bucket_bytes:
            self._launch()

    def function_done(self):
        """End of one Function's backward: launch the partial bucket and make the compute stream wait for this Function's
        reductions (the tensors are about to be accumulated into .grad)."""
        if self.world == 1:
            return
        self._launch()
        if self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)
            self._done = None
        for work, flat, grads in self.inflight:
            work.wait()
            flat.div_(self.world)
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(f)
        self.inflight = []

    def _launch(self):
        grads, self.pending, self.pending_bytes = self.pending, [], 0
        if not grads:
            return
        dist = self.dist
        self.bytes_reduced += sum(g.numel() * g.element_size() for g in grads)
        self.buckets_launched += 1
        if grads[0].is_cuda:
            dev = grads[0].device
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=dev)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))            # the gradients of this bucket are final
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(ready)
                with dist._coalescing_manager(group=self.group, device=dev):
                    for g in grads:
                        dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
                self._done = torch.cuda.Event()
                self._done.record(self._comm)
                for g in grads:
                    g.record_stream(self._comm)
            if self.debug_sync:
                self._comm.synchronize()
        else:
            flat = torch._utils._flatten_dense_tensors(grads)
            self.inflight.append((dist.all_reduce(flat, async_op=True, group=self.group), flat, grads))

    # -- user side -------------------------------------------------------------------------------------------------
    def __enter__(self):
        from . import grad
        self._prev, grad._grad_sink = grad._grad_sink, self
        self.bytes_reduced, self.buckets_launched, self.ids = 0, 0, set()
        return self

    def __exit__(self, *exc):
        from . import grad
        grad._grad_sink = self._prev
        self.function_done()
        return False

    def reduce_rest(self, params):
        """Plain all-reduce of the gradients no plan emitted (parameters of modules that are not training plans)."""
        return allreduce_gradients(list(params), self.dist, self.world)


_reducer_group = None


def reducer_group(dist=None):
    """A process group of its own for the in-backward reducer (NCCL only; None = use the default group).  The bucket
    all-reduces run WHILE the backward kernels do: an NCCL kernel that spreads over many SMs takes them away from conv
    grids sized to fill the machine, so the communicator can be capped (HFC_REDUCER_MAX_CTAS, NCCL's `max_ctas`) and its
    stream raised in priority (HFC_REDUCER_HIGH_PRIORITY=1) so that the few CTAs it does use are scheduled at once."""
    global _reducer_group
    if dist is None:
        import torch.distributed as dist
    if _reducer_group is not None:
        return _reducer_group
    max_ctas = int(os.environ.get("HFC_REDUCER_MAX_CTAS", "0"))
    high = os.environ.get("HFC_REDUCER_HIGH_PRIORITY", "0") == "1"
    if not dist.is_initialized() or dist.get_backend() != "nccl" or (max_ctas <= 0 and not high):
        return None
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=high)
    if max_ctas > 0:
        opts.config.max_ctas = max_ctas
        opts.config.min_ctas = min(max_ctas, 1)
    _reducer_group = dist.new_group(backend="nccl", pg_options=opts)
    return _reducer_group
