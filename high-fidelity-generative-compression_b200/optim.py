"""Adam with torch.optim.Adam's semantics (the optimizer train.py:287-300 constructs) as ONE libhfc launch per
step over all parameters of the group instead of torch's ~80 multi-tensor launches.

    from hific_b200.optim import Adam          # drop-in for torch.optim.Adam(params, lr=...)

Supported options: lr, betas, eps, weight_decay (L2, as torch.optim.Adam); amsgrad / maximize / capturable are not
on the reference's path and raise.  State keys ('step', 'exp_avg', 'exp_avg_sq') match torch's, so state_dicts are
interchangeable with torch.optim.Adam checkpoints (train.py:31-47 saves them)."""
import torch

from ._lib import check, lib
from .ops import _ptr, _stream


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("hific_b200.optim.Adam: amsgrad / maximize are not on the HiFIC path")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._chunk = int(lib.hfc_adam_chunk())
        self._maps = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            self.step_subset(group, [p for p in group["params"] if p.grad is not None])
        return loss

    @torch.no_grad()
    def step_subset(self, group, ps):
        """One Adam update of the parameters `ps` (a subset of `group["params"]`, each with a gradient) -- `step()` is this
        over every group.  Lets a caller step bucket by bucket as gradient all-reduces complete
        (hific_b200.dist.allreduce_then_step); every parameter must be passed exactly once per optimizer step."""
        if not ps:
            return
        dev = ps[0].device
        if not ps[0].is_cuda:
            raise RuntimeError("hific_b200.optim.Adam has no CPU path")
        by_step = {}            # parameters that skipped steps (no gradient) carry their own bias correction
        keep = []               # contiguous copies of strided gradients must outlive the launch that reads them
        for p in ps:
            if p.grad.is_sparse or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("hific_b200.optim.Adam: dense contiguous float32 parameters only")
            st = self.state[p]
            if not st:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            g = p.grad
            if not g.is_contiguous():
                g = g.contiguous()
                keep.append(g)
            by_step.setdefault(int(st["step"]), []).append(
                (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()))
        b1, b2 = group["betas"]
        for step, rows in by_step.items():      # one launch per distinct step count (normally exactly one)
            sizes = tuple(r[4] for r in rows)
            bm = self._maps.get((sizes, dev))
            if bm is None:      # block -> (tensor, chunk) map depends on the sizes only
                pairs = [(t, c) for t, n in enumerate(sizes) for c in range((n + self._chunk - 1) // self._chunk)]
                bm = self._maps[(sizes, dev)] = torch.tensor(pairs, dtype=torch.int32).reshape(-1).to(dev)
            table = torch.tensor(rows, dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
            check(lib.hfc_adam_multi(_ptr(table), _ptr(bm), bm.numel() // 2, float(group["lr"]), float(b1), float(b2),
                                     float(group["eps"]), float(group["weight_decay"]), step, _stream()), "adam_multi")
        # the kernel wrote the parameters behind torch's back: bump their version counters so that autograd's
        # saved-tensor checks and the packed-weight caches (ops.Conv.packed_weights) see the update
        for p in ps:
            torch.autograd.graph.increment_version(p)
        for g in keep:          # the caching allocator must not hand the copies out before the kernel has read them
            g.record_stream(torch.cuda.current_stream(dev))
        del keep
