"""Import the UNMODIFIED reference (read-only, /root/reference) in this container -- TEST INFRASTRUCTURE.

Only usable where /root/reference exists (the build container); nothing that runs on the GPU box may import
this.  Three workarounds, none touching hot-path arithmetic:
  1. `autograd` (HIPS) is not installed: stubbed (only the host ANS coder uses it);
  2. `skimage` is not installed: stubbed (imported by LPIPS / datasets, unused on the path);
  3. no network: torchvision's `alexnet(pretrained=True)` becomes a seeded random trunk.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("HIFIC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src"))


def install():
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    import numpy as np
    if "autograd" not in sys.modules:
        ag = types.ModuleType("autograd")
        ag.numpy = np
        ag.make_vjp = lambda *a, **k: None
        ext = types.ModuleType("autograd.extend")
        ext.primitive = lambda f: f
        ext.defvjp = lambda *a, **k: None
        ext.vspace = lambda *a, **k: None
        ext.VSpace = object
        ag.extend = ext
        sys.modules["autograd"] = ag
        sys.modules["autograd.numpy"] = np
        sys.modules["autograd.extend"] = ext
    for name in ("skimage", "skimage.measure", "skimage.color", "skimage.transform", "skimage.io"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.compare_ssim = m.structural_similarity = m.imread = lambda *a, **k: None
            sys.modules[name] = m
    import torch
    import torchvision

    orig_alexnet = torchvision.models.alexnet

    def alexnet_offline(pretrained=False, **kw):
        state = torch.random.get_rng_state()
        torch.manual_seed(1234)
        net = orig_alexnet(weights=None)
        torch.random.set_rng_state(state)
        return net

    torchvision.models.alexnet = alexnet_offline
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


class NoiseFeeder:
    """Replaces torch.nn.init.uniform_ while active so the reference consumes OUR noise tensors, in order
    (hyper-latents first, then latents -- src/hyperprior.py:65 called from :284 and :305)."""

    def __init__(self, noises):
        self.noises = list(noises)
        self.calls = 0

    def __enter__(self):
        import torch
        self._orig = torch.nn.init.uniform_

        def fake(t, a=0.0, b=1.0):
            n = self.noises[self.calls]
            self.calls += 1
            assert tuple(n.shape) == tuple(t.shape), (n.shape, t.shape)
            with torch.no_grad():
                t.copy_(n)
            return t

        torch.nn.init.uniform_ = fake
        return self

    def __exit__(self, *exc):
        import torch
        torch.nn.init.uniform_ = self._orig
