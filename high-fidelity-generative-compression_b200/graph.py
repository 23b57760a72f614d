"""CUDA-graph capture of a fixed-shape forward pass.

A forward is ~150 libhfc launches plus a handful of scalar torch kernels; at a few milliseconds per step the
Python/ctypes launch path (~15 us per launch) would otherwise bound the step.  Every intermediate buffer is
owned by a plan (hific_b200.engine) and tensor maps travel by value in the kernel parameters, so the whole
sequence is capturable: one graph per (input shape, module state).
"""
import torch


class GraphedCall:
    """Captures ``fn(*static_inputs)`` once; ``__call__`` copies new inputs in and replays."""

    def __init__(self, fn, example_inputs, warmup=2):
        self.static_in = [t.detach().clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):          # builds plans, packs weights, sets kernel attributes
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
