"""Base class of the table-driven entropy models (src/compression/entropy_models.py:21-84)."""
import abc

import numpy as np
import torch
import torch.nn as nn

from . import entropy_coding

MIN_SCALE = 0.11
MIN_LIKELIHOOD = 1e-9
MAX_LIKELIHOOD = 1e4
TAIL_MASS = 2 ** (-8)
PRECISION_P = 16  # precision of the rANS coder


class ContinuousEntropyModel(nn.Module, metaclass=abc.ABCMeta):
    def __init__(self, distribution, likelihood_bound=MIN_LIKELIHOOD, tail_mass=TAIL_MASS, precision=PRECISION_P):
        super().__init__()
        self.distribution = distribution
        self.likelihood_bound = float(likelihood_bound)
        self.tail_mass = float(tail_mass)
        self.precision = int(precision)
        self._host_tables, self._host_tables_key = None, None

    def quantize_st(self, inputs, offsets=None):
        """Straight-through rounding about `offsets` (entropy_models.py:49-63): forward floor(v + .5), identity gradient.
        Plain tensor ops -- API parity only; the hot path rounds inside hfc_latent_likelihood / hfc_quantize_symbols."""
        shift = 0 if offsets is None else offsets.to(inputs)
        centred = inputs - shift
        rounded = centred + (torch.floor(centred + 0.5) - centred).detach()
        return rounded if offsets is None else rounded + shift

    def dequantize(self, x, offsets=None):
        """Symbols -> latents (entropy_models.py:65-73)."""
        return x.to(torch.float32) if offsets is None else x.type_as(offsets) + offsets

    @abc.abstractmethod
    def build_tables(self, **kwargs):
        pass

    def _register_tables(self, cdf, cdf_offset, cdf_length):
        """CDF / CDF_offset / CDF_length as frozen int32 parameters (prior_model.py:109-120): they travel in the
        state_dict, which is how the decoder gets the encoder's exact tables."""
        device = self.CDF.device if hasattr(self, "CDF") else None
        for name, value in (("CDF", cdf), ("CDF_offset", cdf_offset), ("CDF_length", cdf_length)):
            t = torch.as_tensor(np.asarray(value), dtype=torch.int32)
            if device is not None:
                t = t.to(device)
            if hasattr(self, name):
                delattr(self, name)
            self.register_parameter(name, nn.Parameter(t, requires_grad=False))
        self._host_tables = None

    def host_tables(self):
        """numpy copy of the tables for the host coder, refreshed when the parameters change (load_state_dict)."""
        key = tuple((t.data_ptr(), t._version) for t in (self.CDF, self.CDF_length, self.CDF_offset))
        if self._host_tables is None or self._host_tables_key != key:
            self._host_tables = entropy_coding.Tables(self.CDF.detach().cpu().numpy(),
                                                      self.CDF_length.detach().cpu().numpy(),
                                                      self.CDF_offset.detach().cpu().numpy())
            self._host_tables_key = key
        return self._host_tables
