"""hific_b200 -- B200-native (sm_100a) implementation of the HiFIC forward/backward hot path.

The directory is called ``high-fidelity-generative-compression_b200`` (not importable as written);
``import hific_b200`` (the alias package next to it) resolves to this directory.

Layout
    csrc/        hand-written CUDA (tcgen05 / TMA implicit-GEMM conv, fused elementwise) + C ABI
    _lib.py      ctypes loader for libhfc.so (fails loudly when the library is missing)
    ops.py       thin host wrappers: torch tensors in, raw pointers across the C ABI
    network/, normalisation/, hyperprior.py, model.py
                 host-side mirror of the reference's module API (same class names, constructor
                 arguments and state_dict keys as Justin-Tan/high-fidelity-generative-compression)
"""
__version__ = "0.1.0"
