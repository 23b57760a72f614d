"""Multi-GPU plumbing of the training step (one process per GPU, torch.distributed): the HiFIC path is purely
data-parallel (ChannelNorm is per pixel, no batch statistics), so ranks only exchange GRADIENTS -- one coalesced
all-reduce of the stepped parameter group after backward (train.py has no DDP wrapper; north_star asks for "NCCL
allreduce over NVLink for gradients only").  Backend-agnostic: NCCL on the GPUs, gloo in the CPU tests."""
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


def max_over_ranks(value, device, dist=None):
    """The contract's timing rule: a step takes as long as its slowest rank."""
    if dist is None:
        import torch.distributed as dist
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
