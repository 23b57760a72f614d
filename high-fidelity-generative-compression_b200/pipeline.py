"""Host <-> device streaming around `Model.forward` (EVALUATION mode): the H2D copy of batch i+1 and the D2H copy of
batch i-1 run on their own CUDA streams while the kernels of batch i execute, so a serving / evaluation loop
(compress.py:88-113 iterates a DataLoader and moves every batch to the device synchronously) is bound by
max(copy, compute) instead of their sum.  Inputs must be PINNED host tensors; results land in pinned host buffers
owned by the pipeline slot and stay valid until the slot is reused `depth` submissions later.

    pipe = PipelinedForward(model, depth=2)
    t0 = pipe.submit(batch0)               # returns immediately
    t1 = pipe.submit(batch1)
    recon0, q_bpp0 = pipe.result(t0)       # waits for batch 0 only
"""
import torch


class _Slot:
    def __init__(self):
        self.x_dev = self.recon_dev = self.bpp_dev = self.recon_host = self.bpp_host = None
        self.ev_in, self.ev_cmp, self.ev_out = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.used = False


class PipelinedForward:
    def __init__(self, model, depth=2):
        self.model = model
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("PipelinedForward: the model must live on a CUDA device (hific_b200 has no CPU path)")
        self.h2d, self.d2h = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.slots = [_Slot() for _ in range(max(2, depth))]
        self.count = 0

    def submit(self, x_host):
        if not x_host.is_pinned():
            raise ValueError("PipelinedForward.submit: pass a pinned host tensor (x.pin_memory())")
        s = self.slots[self.count % len(self.slots)]
        ticket = self.count
        self.count += 1
        main = torch.cuda.current_stream(self.dev)
        if s.x_dev is None or s.x_dev.shape != x_host.shape:
            s.x_dev = torch.empty(x_host.shape, dtype=x_host.dtype, device=self.dev)
            s.recon_dev = s.recon_host = None
            # The block comes from the MAIN stream's pool: it may be memory that kernels still queued on the main stream
            # (temporaries of the previous submission, already freed on the host) have yet to write.  The copy stream
            # must not touch it before they are done -- without this wait the second submission's input was overwritten
            # by the tail of the first forward (intermittent in round 1, every time once the forward got faster).
            self.h2d.wait_stream(main)
        with torch.cuda.stream(self.h2d):
            if s.used:
                self.h2d.wait_event(s.ev_cmp)        # the kernels that read this slot's input have finished
            s.x_dev.copy_(x_host, non_blocking=True)
            s.ev_in.record(self.h2d)
        main.wait_event(s.ev_in)
        if s.used:
            main.wait_event(s.ev_out)                # the previous D2H out of this slot's staging buffers is done
        with torch.no_grad():
            recon, q_bpp = self.model(s.x_dev, writeout=False)
            if s.recon_dev is None:
                s.recon_dev, s.bpp_dev = torch.empty_like(recon), torch.empty_like(q_bpp, dtype=torch.float32)
                s.recon_host = torch.empty(recon.shape, dtype=recon.dtype).pin_memory()
                s.bpp_host = torch.empty(q_bpp.shape, dtype=torch.float32).pin_memory()
            s.recon_dev.copy_(recon)                 # staging: the model's own output buffers are reused by the next call
            s.bpp_dev.copy_(q_bpp)
        s.ev_cmp.record(main)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s.ev_cmp)
            s.recon_host.copy_(s.recon_dev, non_blocking=True)
            s.bpp_host.copy_(s.bpp_dev, non_blocking=True)
            s.ev_out.record(self.d2h)
        s.used = True
        return ticket

    def result(self, ticket):
        """(reconstruction, q_bpp) of submission `ticket` as pinned host tensors (valid until the slot is reused)."""
        if ticket < self.count - len(self.slots) or ticket >= self.count:
            raise ValueError("PipelinedForward.result: the ticket's slot has been reused or was never submitted")
        s = self.slots[ticket % len(self.slots)]
        s.ev_out.synchronize()
        return s.recon_host, s.bpp_host

    def drain(self):
        """Make the current stream wait for every outstanding copy (so that an event recorded next covers them)."""
        main = torch.cuda.current_stream(self.dev)
        for s in self.slots:
            if s.used:
                main.wait_event(s.ev_out)
