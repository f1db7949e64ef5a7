"""Drop-in replacement for the reference's MixStyle / DSU module (src/advanced/mixstyle.py:6-108), the sibling of
MaxStyle used by the baseline configurations and by `generate_style_augmented_latent_code`
(src/models/advanced_triplet_recon_segmentation_model.py:632-670).  SURVEY.md section 8f-2.

Same class name, constructor, `forward(x, perm=None)`, `update_mix_method`, `get_perm`, `__repr__` and generator
consumption order (CPU: rand(1) -> mixing weight -> randperm; device: randn x2 for mix='gaussian'), so the same seed gives
the same augmentation.  The arithmetic is the MaxStyle kernels with a different table rule:
  mix in {'random','crossdomain'}:  MIX_STYLE | NO_NOISE | NO_CLAMP   (lerp with the partner, weight used as drawn)
  mix == 'gaussian' (DSU):          noise only, eps_sig / eps_mu ~ N(0,1) scaled by the batch std, recomputed every call
eps defaults to 1e-8 here (1e-6 in MaxStyle).  No learnable parameters; dX flows through the same backward kernel.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import functional as F


class _CallState:
    """What MaxStyleFunction needs from a 'layer' for ONE call (MixStyle keeps no state between calls)."""
    _fused_step = None

    def __init__(self, flags: int, eps: float, num_feature: int, perm_dev, workspace):
        self._f, self.eps, self.num_feature = flags, eps, num_feature
        self._perm, self._ws = perm_dev, workspace
        self.gamma_std = None
        self.beta_std = None
        self.gamma_noise = self.beta_noise = self.lmda = None

    def _flags(self):
        return self._f

    def _perm_device(self, device):
        return self._perm

    def _workspace_for(self, x):
        return self._ws


class MixStyle(nn.Module):
    """MixStyle (Zhou et al., ICLR 2021) and DSU-style Gaussian statistics noise on B200 kernels.

    Args mirror the reference (mixstyle.py:14-22): p, alpha (Beta parameter), eps, mix in
    {'random','crossdomain','gaussian'}, lmda (fixed weight; values outside [0,1] extrapolate).
    As in the reference the constructor ignores `coefficient_sampler`; set the attribute `coeficient_sampler`
    to 'beta' | 'uniform' | 'gaussian' afterwards to change the sampler."""

    def __init__(self, p=0.5, alpha=0.1, eps=1e-8, mix='random', lmda=None, zero_init=False, coefficient_sampler=None):
        super().__init__()
        self.p = p
        self.eps = eps
        self.mu = None
        self.std = None
        self.zero_init = zero_init
        self.alpha = alpha
        self.mix = mix
        self._activated = True
        self.lmda = lmda
        self.coeficient_sampler = None
        self.beta = torch.distributions.Beta(alpha, alpha)
        self._workspace: Optional[torch.Tensor] = None
        self._workspace_key = None

    def __repr__(self):
        return f'MixStyle(p={self.p}, alpha={self.alpha}, eps={self.eps}, mix={self.mix})'

    def update_mix_method(self, mix='random'):
        self.mix = mix

    def get_perm(self):
        return self.perm

    def _workspace_for(self, x):
        layout = F.layout_of(x)
        key = (x.device, tuple(x.shape), x.dtype, layout)
        if self._workspace_key != key:
            n, c, h, w = x.shape
            self._workspace = F.new_workspace(n, c, h, w, F.dtype_code(x), x.device, layout)
            self._workspace_key = key
        return self._workspace

    def forward(self, x, perm=None):
        p = torch.rand(1)
        if p > self.p:
            return x
        B, C = x.size(0), x.size(1)
        if self.lmda is None:
            if self.coeficient_sampler is None or self.coeficient_sampler == 'beta':
                lmda = self.beta.sample((B, 1, 1, 1))
            elif self.coeficient_sampler == 'uniform':
                lmda = torch.rand(B, 1, 1, 1)
            elif self.coeficient_sampler == 'gaussian':
                lmda = torch.randn(B, 1, 1, 1)
            else:
                raise ValueError
        else:
            lmda = torch.ones(B, 1, 1, 1) * self.lmda
        if x.dim() != 4:
            raise RuntimeError(f"maxstyle_b200: expected a 4-d [N,C,H,W] feature map, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("maxstyle_b200: MixStyle.forward got a CPU tensor; the layer runs only as CUDA kernels "
                               "on a B200 and has no CPU fallback")
        if x.size(2) * x.size(3) < 2:
            raise RuntimeError("maxstyle_b200: MixStyle needs at least 2 elements per plane (unbiased variance)")
        lmda = lmda.to(device=x.device, dtype=torch.float32)
        with F.device_guard(x.device):
            if self.mix in ('random', 'crossdomain'):
                if perm is None:
                    if self.mix == 'random':
                        perm = torch.randperm(B)
                    else:
                        perm = torch.arange(B - 1, -1, -1)
                        perm_b, perm_a = perm.chunk(2)
                        perm_b = perm_b[torch.randperm(B // 2)]
                        perm_a = perm_a[torch.randperm(B // 2)]
                        perm = torch.cat([perm_b, perm_a], 0)
                self.perm = perm
                perm_t = torch.as_tensor(perm)
                # range check on the HOST copy only (the reference's perm is a CPU tensor): no device synchronisation on the
                # forward path, so the call stays graph-capturable; a caller-supplied CUDA perm is taken as it is
                if perm_t.numel() != B or (not perm_t.is_cuda and B > 0 and (int(perm_t.min()) < 0 or int(perm_t.max()) >= B)):
                    raise IndexError("maxstyle_b200: perm must hold B indices into the batch")
                perm_dev = perm_t.to(device=x.device, dtype=torch.int64)
                state = _CallState(L.FLAG_MIX_STYLE | L.FLAG_NO_NOISE | L.FLAG_NO_CLAMP, self.eps, C, perm_dev,
                                   self._workspace_for(F.dense_layout(x)))
                state.gamma_std = torch.zeros(1, C, 1, 1, device=x.device)        # unused with NO_NOISE; keeps the one-kernel path
                state.beta_std = torch.zeros(1, C, 1, 1, device=x.device)
                return F.MaxStyleFunction.apply(x, None, None, lmda.reshape(B), state, L.PRE_NONE, 0.0, None)
            elif self.mix == 'gaussian':
                gaussian_mu = torch.randn(B, C, 1, 1, device=x.device)            # scaled by std_n(mu)  -> beta-like noise
                gaussian_std = torch.randn(B, C, 1, 1, device=x.device)           # scaled by std_n(sig) -> gamma-like noise
                state = _CallState(0, self.eps, C, None, self._workspace_for(F.dense_layout(x)))
                # gamma_std / beta_std stay None: the batch std is taken inside the kernel on every call (:101-102)
                return F.MaxStyleFunction.apply(x, gaussian_std, gaussian_mu, None, state, L.PRE_NONE, 0.0, None)
            else:
                raise NotImplementedError
