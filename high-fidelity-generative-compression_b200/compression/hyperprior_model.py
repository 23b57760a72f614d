"""Factorized hyper-latent density (parameter contract of the reference's HyperpriorDensity,
src/compression/hyperprior_model.py:252-387).  The likelihood itself is evaluated by the fused kernel
`hfc_hyperlatent_likelihood` (see hific_b200.hyperprior.Hyperprior.forward); this module owns the
parameters H_k, a_k, b_k and their packed (softplus / tanh pre-applied) device copy.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops

MIN_LIKELIHOOD = 1e-9
MAX_LIKELIHOOD = 1e3


class HyperpriorDensity(nn.Module):
    def __init__(self, n_channels, init_scale=10., filters=(3, 3, 3), min_likelihood=MIN_LIKELIHOOD,
                 max_likelihood=MAX_LIKELIHOOD, **kwargs):
        super().__init__()
        if tuple(filters) != (3, 3, 3):
            raise NotImplementedError("the fused density kernel is built for filters=(3, 3, 3)")
        self.init_scale = float(init_scale)
        self.filters = tuple(int(f) for f in filters)
        self.min_likelihood, self.max_likelihood = float(min_likelihood), float(max_likelihood)
        self.n_channels = n_channels
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for k in range(len(self.filters) + 1):
            H = nn.Parameter(torch.full((n_channels, f[k + 1], f[k]), float(np.log(np.expm1(1 / scale / f[k + 1])))))
            a = nn.Parameter(torch.zeros((n_channels, f[k + 1], 1)))
            b = nn.Parameter(torch.zeros((n_channels, f[k + 1], 1)))
            torch.nn.init.uniform_(b, -0.5, 0.5)
            self.register_parameter(f"H_{k}", H)
            self.register_parameter(f"a_{k}", a)
            self.register_parameter(f"b_{k}", b)
        self._packed, self._packed_key = None, None

    def _tensors(self):
        return ([getattr(self, f"H_{k}") for k in range(4)], [getattr(self, f"a_{k}") for k in range(4)],
                [getattr(self, f"b_{k}") for k in range(4)])

    def packed_params(self):
        Hs, a_s, bs = self._tensors()
        key = tuple((t.data_ptr(), t._version) for t in Hs + a_s + bs)
        if self._packed is None or key != self._packed_key:
            self._packed, self._packed_key = ops.pack_density_params(Hs, a_s, bs), key
        return self._packed

    def forward(self, x, **kwargs):
        raise NotImplementedError(
            "HyperpriorDensity is evaluated inside Hyperprior.forward by the fused likelihood kernel; the "
            "per-element likelihood tensor is never materialised on the B200 path")
