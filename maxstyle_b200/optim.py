"""Optimiser step fused into the backward epilogue (BASELINE.json north_star item 4).

The reference caller builds `torch.optim.Adam(style_layers.parameters(), lr=0.1)` and calls
`loss.backward(); optimizer.step()` (advanced_triplet_recon_segmentation_model.py:537,561-562).
`FusedStyleOptimizer` has the same constructor/step/zero_grad surface, but the update is applied
by the backward kernel's per-sample epilogue, so `step()` launches nothing.
"""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib as L
from .functional import FusedStepState, StepConfig
from .layer import MaxStyle


class FusedStyleOptimizer:
    """Adam (reference semantics) or sign-gradient step executed inside MaxStyle's backward.

    Args:
        layers: MaxStyle modules (e.g. `nn.ModuleDict.values()`); inactive layers are skipped.
        lr: step size (reference: 0.1).
        mode: 'adam' -- torch.optim.Adam defaults, what the reference uses on loss = -CE;
              'sign' -- p <- p - lr*sign(grad)  (with maximize=True: p <- p + lr*sign(grad)).
        maximize: ascend the back-propagated loss instead of descending it.
        keep_grads: also materialise .grad on the parameters (costs three small stores).
    """

    def __init__(self, layers: Iterable[MaxStyle], lr: float = 0.1, mode: str = "adam", betas=(0.9, 0.999),
                 eps: float = 1e-8, maximize: bool = False, keep_grads: bool = False):
        modes = {"adam": L.STEP_ADAM, "sign": L.STEP_SIGN}
        if mode not in modes:
            raise ValueError(f"mode must be one of {sorted(modes)}, got {mode!r}")
        self.cfg = StepConfig(mode=modes[mode], lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, maximize=maximize)
        self.layers = []
        for layer in layers:
            if not isinstance(layer, MaxStyle):
                raise TypeError(f"expected MaxStyle modules, got {type(layer).__name__}")
            learn_noise = isinstance(layer.gamma_noise, torch.nn.Parameter) and layer.gamma_noise.requires_grad
            learn_mix = isinstance(layer.lmda, torch.nn.Parameter) and layer.lmda.requires_grad
            if not (learn_noise or learn_mix):
                continue
            st = FusedStepState(self.cfg, layer.gamma_noise, layer.beta_noise, layer.lmda, learn_noise, learn_mix)
            st.keep_grads = keep_grads
            layer._fused_step = st
            self.layers.append(layer)

    def step(self):
        """No-op: the update already happened in the backward epilogue of each layer."""
        return None

    def zero_grad(self, set_to_none: bool = True):
        for layer in self.layers:
            layer.zero_grad(set_to_none=set_to_none)

    def detach(self):
        """Stop stepping: later backward passes only produce gradients again."""
        for layer in self.layers:
            layer._fused_step = None
        self.layers = []

    def step_count(self, layer: MaxStyle) -> int:
        return int(layer._fused_step.step_dev.item())
