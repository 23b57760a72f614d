"""Hyperprior analysis / synthesis networks (reference: src/network/hyper.py:45-97)."""
import torch
import torch.nn as nn

from .. import engine, train_plan


class HyperpriorAnalysis(nn.Module):
    def __init__(self, C=220, N=320, activation='relu'):
        super().__init__()
        if activation != 'relu':
            raise NotImplementedError("only ReLU is built")
        self.C, self.N = C, N
        self.n_downsampling_layers = 2
        self.conv1 = nn.Conv2d(C, N, kernel_size=3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(N, N, kernel_size=5, stride=2, padding=2, padding_mode='reflect')
        self.conv3 = nn.Conv2d(N, N, kernel_size=5, stride=2, padding=2, padding_mode='reflect')
        self._plans = engine.PlanCache(lambda y: engine.HyperAnalysisPlan(y.shape[0], y.shape[2], y.shape[3],
                                                                            self.C, self.N, y.device))
        self._train_plans = engine.PlanCache(lambda y: train_plan.HyperAnalysisTrainPlan(
            y.shape[0], y.shape[2], y.shape[3], self.C, self.N, y.device))

    def _apply(self, fn, *a, **k):
        self._plans.clear()
        self._train_plans.clear()
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        engine._require_cuda(x, "HyperpriorAnalysis")
        if engine.wants_grad(self, x):
            return train_plan.run_training(self._train_plans.get(x), x, list(self.parameters()))
        return self._plans.get(x).run(self, x.contiguous())


class HyperpriorSynthesis(nn.Module):
    def __init__(self, C=220, N=320, activation='relu', final_activation=None):
        super().__init__()
        if activation != 'relu' or final_activation is not None:
            raise NotImplementedError("only ReLU / no final activation is built (the HiFIC default)")
        self.C, self.N = C, N
        self.final_activation = None
        self.conv1 = nn.ConvTranspose2d(N, N, kernel_size=5, stride=2, padding=2, output_padding=1)
        self.conv2 = nn.ConvTranspose2d(N, N, kernel_size=5, stride=2, padding=2, output_padding=1)
        self.conv3 = nn.ConvTranspose2d(N, C, kernel_size=3, stride=1, padding=1)
        self._plans = engine.PlanCache(lambda z: engine.HyperSynthesisPlan(z.shape[0], z.shape[2], z.shape[3],
                                                                             self.C, self.N, z.device))
        self._train_plans = engine.PlanCache(lambda z: train_plan.HyperSynthesisTrainPlan(
            z.shape[0], z.shape[2], z.shape[3], self.C, self.N, z.device))

    def _apply(self, fn, *a, **k):
        self._plans.clear()
        self._train_plans.clear()
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        engine._require_cuda(x, "HyperpriorSynthesis")
        if engine.wants_grad(self, x):
            return train_plan.run_training(self._train_plans.get(x), x, list(self.parameters()))
        return self._plans.get(x).run(self, x.contiguous())


def get_num_DLMM_channels(C, K=4, params=('mu', 'scale', 'mix')):
    """src/network/hyper.py:8-13."""
    return C * K * len(params)


def get_num_mixtures(K_agg, C, params=('mu', 'scale', 'mix')):
    return K_agg // (len(params) * C)


class HyperpriorSynthesisDLMM(nn.Module):
    """Mixture-parameter network of the `-LMM` variant (src/network/hyper.py:100-130): z -> (N, 3*K*C, H, W)."""

    def __init__(self, C=64, N=320, activation='relu', final_activation=None):
        super().__init__()
        if activation != 'relu' or final_activation is not None:
            raise NotImplementedError("only ReLU / no final activation is built (the reference's default)")
        self.C, self.N = C, N
        self.final_activation = None
        self.conv1 = nn.ConvTranspose2d(N, N, kernel_size=5, stride=2, padding=2, output_padding=1)
        self.conv2 = nn.ConvTranspose2d(N, N, kernel_size=5, stride=2, padding=2, output_padding=1)
        self.conv3 = nn.ConvTranspose2d(N, C, kernel_size=3, stride=1, padding=1)
        self.conv_out = nn.Conv2d(C, get_num_DLMM_channels(C), kernel_size=1, stride=1)
        n_out = get_num_DLMM_channels(C)
        self._plans = engine.PlanCache(lambda z: engine.HyperSynthesisDLMMPlan(z.shape[0], z.shape[2], z.shape[3],
                                                                                 self.C, self.N, n_out, z.device))
        self._train_plans = engine.PlanCache(lambda z: train_plan.HyperSynthesisDLMMTrainPlan(
            z.shape[0], z.shape[2], z.shape[3], self.C, self.N, n_out, z.device))

    def _apply(self, fn, *a, **k):
        self._plans.clear()
        self._train_plans.clear()
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        engine._require_cuda(x, "HyperpriorSynthesisDLMM")
        if engine.wants_grad(self, x):
            return train_plan.run_training(self._train_plans.get(x), x, list(self.parameters()))
        return self._plans.get(x).run(self, x.contiguous())
