"""LPIPS (AlexNet, 'net-lin' v0.1) perceptual loss with the reference's call surface
(src/loss/perceptual_similarity/perceptual_loss.py:10-40 -> dist_model.py:105 -> networks_basic.py:61-89).

What runs where: the per-layer "normalise over channels, squared difference, 1x1 lin, spatial mean" reduction is
the fused kernel `hfc_lpips_layer` (5 launches instead of ~50 eager kernels).  The frozen AlexNet trunk is
torchvision's module executed by cuDNN -- SURVEY.md section 8f lists the trunk convs as the first "next" row; they
are not part of this round.  As in the reference, the LPIPS weights are NOT part of Model.state_dict().
"""
import os

import torch
import torch.nn as nn

from .. import ops

_LIN_CHANNELS = (64, 192, 384, 256, 256)
_SLICES = ((0, 2), (2, 5), (5, 8), (8, 10), (10, 12))   # relu1..relu5 of torchvision alexnet.features


def _find_lin_weights():
    cands = [os.environ.get("HIFIC_LPIPS_WEIGHTS", ""),
             os.path.join(os.environ.get("HIFIC_REFERENCE_ROOT", "/root/reference"),
                          "src/loss/perceptual_similarity/weights/v0.1/alex.pth")]
    for c in cands:
        if c and os.path.exists(c):
            return c
    return None


def _synthetic_or_raise(what):
    """The reference always loads pretrained weights (pretrained_networks.py:59, networks_basic.py weights/v0.1/alex.pth).
    Training against a random trunk silently optimises k_P * noise, so a missing checkpoint is an ERROR unless the caller
    opts in to the seeded stand-in (tests, bench and smoke on the network-less GPU box do: HFC_LPIPS_SYNTHETIC=1)."""
    if os.environ.get("HFC_LPIPS_SYNTHETIC") != "1":
        raise RuntimeError(f"PerceptualLoss: {what}.  Provide the checkpoint, pass trunk= / lin_weights=, or set "
                           "HFC_LPIPS_SYNTHETIC=1 to accept a seeded random stand-in (parity tests / benchmarks only).")
    import warnings
    warnings.warn(f"PerceptualLoss: {what}; using the seeded synthetic stand-in (HFC_LPIPS_SYNTHETIC=1)", stacklevel=3)


class PerceptualLoss(nn.Module):
    def __init__(self, model='net-lin', net='alex', colorspace='rgb', spatial=False, use_gpu=True, gpu_ids=[0],
                 version='0.1', trunk=None, lin_weights=None):
        super().__init__()
        if model != 'net-lin' or net != 'alex' or spatial or version != '0.1':
            raise NotImplementedError("only LPIPS net-lin / alex / v0.1 (what HiFIC uses) is built")
        if trunk is None:
            import torchvision
            cached = os.path.join(torch.hub.get_dir(), "checkpoints", "alexnet-owt-7be5be79.pth")
            if os.path.exists(cached):   # ImageNet weights only if already cached: never touch the network
                trunk = torchvision.models.alexnet(weights="IMAGENET1K_V1").features
            else:                        # offline stand-in (same seed as oracle/ref_shim.py)
                _synthetic_or_raise("the ImageNet AlexNet checkpoint (alexnet-owt-7be5be79.pth) is not in the torch hub cache "
                                    f"({cached}): a randomly initialised trunk is NOT LPIPS")
                self._synthetic_trunk = True
                state = torch.random.get_rng_state()
                torch.manual_seed(1234)
                trunk = torchvision.models.alexnet(weights=None).features
                torch.random.set_rng_state(state)
        self.trunk = trunk.eval()
        for p in self.trunk.parameters():
            p.requires_grad = False
        self.register_buffer('shift', torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('scale', torch.tensor([.458, .448, .450])[None, :, None, None])
        lins = []
        path, sd = None, (lin_weights or {})
        if lin_weights is None:
            # the 1 152 'lin' weights of LPIPS v0.1 / alex (BSD-licensed data published with LPIPS, vendored by the
            # reference under weights/v0.1/alex.pth) ship with the package as a plain .npz
            packaged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights",
                                    "lpips_alex_lin_v0.1.npz")
            if os.path.exists(packaged):
                import numpy as np
                z = np.load(packaged)
                sd = {f"lin{k}.model.1.weight": torch.from_numpy(z[f"lin{k}"]) for k in range(5)}
                path = packaged
            else:
                path = _find_lin_weights()
                sd = torch.load(path, map_location='cpu') if path else {}
        for k, c in enumerate(_LIN_CHANNELS):
            key = f"lin{k}.model.1.weight"
            if key in sd:
                w = sd[key].reshape(-1).float()
            else:   # deterministic non-negative stand-in when the vendored file is not reachable
                _synthetic_or_raise(f"LPIPS 'lin' weights ({key}) not found")
                g = torch.Generator().manual_seed(100 + k)
                w = torch.rand(c, generator=g) * 0.02
            lins.append(nn.Parameter(w.clone(), requires_grad=False))
        self.lins = nn.ParameterList(lins)
        self.lin_source = path or ("provided" if lin_weights else "synthetic")
        self.trunk_source = "provided / ImageNet checkpoint" if not getattr(self, "_synthetic_trunk", False) else "synthetic"

    def features(self, x):
        h = (x - self.shift) / self.scale                      # ScalingLayer, networks_basic.py:91-98
        outs = []
        for lo, hi in _SLICES:
            for i in range(lo, hi):
                h = self.trunk[i](h)
            outs.append(h)
        return outs

    def native_trunk(self):
        """The AlexNet trunk runs on the tcgen05 conv kernel (loss/lpips_trunk.py) by default -- 27.2 vs 29.2 ms per c2
        training step on the B200 (round 2, gpurun_out/c2_bench_full.json), gradients as close to the fp32 trunk as
        cuDNN's own TF32 path (tests/test_gpu_zzlpips_trunk.py).  HFC_LPIPS_TRUNK=cudnn selects torchvision's module
        executed by cuDNN + torch autograd."""
        return os.environ.get("HFC_LPIPS_TRUNK", "native") == "native"

    def _native(self, pred, target, normalize):
        from .. import engine
        from . import lpips_trunk
        if getattr(self, "_native_plans", None) is None:
            self._native_plans = engine.PlanCache(
                lambda x: lpips_trunk.LpipsTrunkPlan(x.shape[0], x.shape[2], x.shape[3], x.device))
        plan = self._native_plans.get(pred)
        if torch.is_grad_enabled() and pred.requires_grad:
            return lpips_trunk.LpipsTrunkFn.apply(pred, target.detach(), plan, self, bool(normalize)).view(-1, 1, 1, 1)
        with torch.no_grad():
            return plan.forward(self, target, pred, bool(normalize)).view(-1, 1, 1, 1)

    def forward(self, pred, target, normalize=False):
        """Returns (N, 1, 1, 1) like the reference."""
        if not pred.is_cuda:
            raise RuntimeError("PerceptualLoss: hific_b200 has no CPU path")
        if self.native_trunk():
            return self._native(pred, target, normalize)
        if normalize:
            target = 2 * target - 1
            pred = 2 * pred - 1
        if torch.is_grad_enabled() and pred.requires_grad:
            # training: torch autograd carries the gradient through the frozen cuDNN trunk; the per-layer feature
            # loss and its gradient w.r.t. the reconstruction's features are the fused kernels
            with torch.no_grad():
                f0 = self.features(target)
            f1 = self.features(pred)
            out = 0
            for k in range(5):
                out = out + ops.LpipsLayerFn.apply(f0[k], f1[k], self.lins[k])
            return out.view(-1, 1, 1, 1)
        with torch.no_grad():
            f0, f1 = self.features(target), self.features(pred)    # model.forward(target, pred), perceptual_loss.py:40
            out = torch.zeros(pred.shape[0], dtype=torch.float32, device=pred.device)
            for k in range(5):
                ops.lpips_layer(f0[k], f1[k], self.lins[k], out)
        return out.view(-1, 1, 1, 1)
