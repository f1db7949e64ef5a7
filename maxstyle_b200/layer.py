"""Drop-in replacement for the reference's MaxStyle module (src/advanced/maxstyle.py:6-189).

Same class name, constructor signature, `forward(x)`, `reset()`, `init_parameters()`,
`__repr__`, parameter names (`gamma_noise`, `beta_noise`, `lmda`, in that order) and public
attributes (`perm`, `rand_p`, `gamma_std`, `beta_std`, `data`, `device`, ...), so that
`MyDecoder.apply_max_style` (src/models/ebm/encoder_decoder.py:598-631) and
`generate_max_style_image` (src/models/advanced_triplet_recon_segmentation_model.py:458-571)
use it unchanged.  The arithmetic runs in libmaxstyle_b200.so (hand-written sm_100a kernels);
the random state is drawn from torch's generators in exactly the reference's call order so
that identical seeds give identical modules.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import functional as F


class MaxStyle(nn.Module):
    """MaxStyle feature-style layer (Chen et al., MICCAI 2022) on B200 kernels.

    Args mirror the reference constructor (maxstyle.py:14-15):
        batch_size, num_feature: N and C of the feature map this instance will see.
        p: probability that the layer is active for this instance (drawn once per init).
        mix_style: lerp each sample's (mu, sig) with those of a batch-permuted partner.
        no_noise: drop the gamma/beta style-noise perturbation.
        mix_learnable / noise_learnable: make `lmda` / `gamma_noise`,`beta_noise` trainable.
        always_use_beta: draw lmda ~ Beta(alpha, alpha) instead of U(0, 1).
        eps: added to the variance before the square root.
        use_gpu: parameters live on 'cuda' (True) or 'cpu'.  CPU instances can be constructed
            (state/RNG parity) but cannot run: forward has no CPU path.
    """

    def __init__(self, batch_size, num_feature, p=0.5, mix_style=True, no_noise=False,
                 mix_learnable=True, noise_learnable=True, always_use_beta=False, alpha=0.1, eps=1e-6,
                 use_gpu=True, debug=False):
        super().__init__()
        self.batch_size = batch_size
        self.num_feature = num_feature
        self.p = p
        self.mix_style = mix_style
        self.no_noise = no_noise
        self.mix_learnable = mix_learnable
        self.noise_learnable = noise_learnable
        self.always_use_beta = always_use_beta
        self.alpha = alpha
        self.eps = eps
        self.use_gpu = use_gpu
        self.debug = debug
        self.device = torch.device("cuda" if use_gpu else "cpu")
        self.data = None
        # replacement-only state (not part of the reference surface)
        self._perm_dev: Optional[torch.Tensor] = None
        self._workspace: Optional[torch.Tensor] = None
        self._workspace_key = None
        self._fused_step = None
        self.init_parameters()

    # ------------------------------------------------------------------------------------
    # random state.  Generator consumption order is the seed contract (maxstyle.py:48-122):
    #   CPU: randperm (redrawn while it is the identity), rand(1);
    #   then only if active: [no_noise: randn x2 on device] -> [noise_learnable: normal_ x2 on
    #   device] -> [mix_style: Beta sample on CPU | rand on device].
    # ------------------------------------------------------------------------------------
    def _draw_permutation(self):
        n = self.batch_size
        identity = torch.arange(n)
        perm = torch.randperm(n)
        # only the full identity is rejected; fixed points are allowed.  (batch_size <= 1 has no
        # non-identity permutation: the reference loops forever there, we keep the identity --
        # such a module can only ever take the B <= 1 early-out.)
        while n > 1 and torch.equal(perm, identity):
            perm = torch.randperm(n)
        return perm

    def _agree_on_draw(self):
        """Hook between the two CPU draws (perm, rand_p) and everything that depends on them.  The multi-GPU layer
        overrides it to take rank 0's draw on every rank BEFORE parameters are allocated from rand_p."""

    def init_parameters(self):
        n, c, dev = self.batch_size, self.num_feature, self.device
        self.perm = self._draw_permutation()
        self._perm_dev = None
        if self.debug:
            print("permutation index", self.perm)
        self.rand_p = torch.rand(1)
        self._agree_on_draw()
        active = bool(self.rand_p < self.p)

        def const_table(shape):
            t = torch.zeros(*shape, device=dev, dtype=torch.float32)
            t.requires_grad = False
            return t

        if not active:
            if self.debug:
                print("not performing maxstyle")
            self.gamma_noise = const_table((n, c, 1, 1))
            self.beta_noise = const_table((n, c, 1, 1))
            self.lmda = const_table((n, 1, 1, 1))
        else:
            if self.no_noise:      # two device draws whose values the forward never uses (reference behaviour)
                fixed_gamma = torch.randn(n, c, 1, 1, device=dev).float()
                fixed_beta = torch.randn(n, c, 1, 1, device=dev).float()
            else:
                fixed_gamma, fixed_beta = const_table((n, c, 1, 1)), const_table((n, c, 1, 1))
            self.gamma_noise = None          # drop earlier registrations so Parameter <-> tensor swaps work
            self.beta_noise = None
            if self.noise_learnable:
                assert self.no_noise is False, "turn no_noise=False to enable the optimization of noise"
                self.gamma_noise = nn.Parameter(torch.empty(n, c, 1, 1, device=dev))
                self.beta_noise = nn.Parameter(torch.empty(n, c, 1, 1, device=dev))
                nn.init.normal_(self.gamma_noise)
                nn.init.normal_(self.beta_noise)
            else:
                fixed_gamma.requires_grad = False
                fixed_beta.requires_grad = False
                self.gamma_noise, self.beta_noise = fixed_gamma, fixed_beta
            self.lmda = None
            if not self.mix_style:
                self.lmda = const_table((n, 1, 1, 1))
            else:
                if self.always_use_beta:
                    self.beta_sampler = torch.distributions.Beta(self.alpha, self.alpha)
                    mix = self.beta_sampler.sample((n, 1, 1, 1)).to(dev)
                else:
                    mix = torch.rand(n, 1, 1, 1, dtype=torch.float32, device=dev)
                self.lmda = nn.Parameter(mix.float())
                self.lmda.requires_grad = bool(self.mix_learnable)
        self.gamma_std = None
        self.beta_std = None
        if self._fused_step is not None:
            self._fused_step = None          # moments belong to the previous parameters
        if self.debug:
            print("lmda:", self.lmda)
            print("gamma_noise:", self.gamma_noise)
            print("beta_noise:", self.beta_noise)
            print("perm:", self.perm)

    def __setattr__(self, name, value):
        # nn.Module refuses to overwrite a registered Parameter with a plain tensor/None;
        # init_parameters() needs exactly that when reset() re-draws an inactive state.
        if name in ("gamma_noise", "beta_noise", "lmda") and not isinstance(value, nn.Parameter):
            self._parameters.pop(name, None)
            object.__setattr__(self, name, value)
            return
        if name in ("gamma_noise", "beta_noise", "lmda"):
            self.__dict__.pop(name, None)
        super().__setattr__(name, value)

    @torch.no_grad()
    def reinit_(self):
        """In-place `reset()` for CUDA-graph replay (StyleLoopExecutor): re-draws perm / rand_p / parameters with exactly the
        generator consumption of `init_parameters()` (so a seeded run matches three fresh constructions in the reference's
        loop, model:522-527) but writes into the EXISTING tensors -- parameter, permutation, batch-std and optimiser-state
        addresses captured in a graph stay valid.  Requires a module built with every learnable it can ever need
        (construct with p=1.0 storage via StyleLoopExecutor); returns whether the layer is active for this draw."""
        n, c, dev = self.batch_size, self.num_feature, self.device
        self.perm = self._draw_permutation()
        self.rand_p = torch.rand(1)
        self._agree_on_draw()
        if self._perm_dev is not None:
            self._perm_dev.copy_(self.perm.to(torch.int64), non_blocking=False)
        active = bool(self.rand_p < self.p)
        if active:
            if self.no_noise:
                torch.randn(n, c, 1, 1, device=dev); torch.randn(n, c, 1, 1, device=dev)     # drawn and unused, as in the reference
            if self.noise_learnable:
                nn.init.normal_(self.gamma_noise)
                nn.init.normal_(self.beta_noise)
            if self.mix_style:
                if self.always_use_beta:
                    mix = torch.distributions.Beta(self.alpha, self.alpha).sample((n, 1, 1, 1)).to(dev)
                else:
                    mix = torch.rand(n, 1, 1, 1, dtype=torch.float32, device=dev)
                self.lmda.copy_(mix.float())
        self._redraw_batch_std = True            # the next forward recomputes gamma_std / beta_std into the same buffers
        st = self._fused_step
        if st is not None:
            for t in (st.gamma_m, st.gamma_v, st.beta_m, st.beta_v, st.lmda_m, st.lmda_v, st.step_dev):
                t.zero_()
        return active

    def reset(self):
        """Re-draw perm / rand_p / parameters and drop the cached batch statistics.
        As in the reference this creates NEW Parameter objects (an optimizer built earlier
        keeps the old ones)."""
        self.init_parameters()
        if self.debug:
            print("reinitializing parameters")

    def __repr__(self):
        if self.p >= self.rand_p:
            return ("MaxStyle: mean of gamma noise: {}, std:{} , mean of beta noise: {}, std: {}, "
                    "mean of mix coefficient: {}, std: {}").format(
                torch.mean(self.gamma_noise), torch.std(self.gamma_noise), torch.mean(self.beta_noise),
                torch.std(self.beta_noise), torch.mean(self.lmda), torch.std(self.lmda))
        return "diffuse style not applied"

    # ------------------------------------------------------------------------------------
    # helpers used by the autograd Function
    # ------------------------------------------------------------------------------------
    def _flags(self) -> int:
        return (L.FLAG_MIX_STYLE if self.mix_style else 0) | (L.FLAG_NO_NOISE if self.no_noise else 0)

    def _perm_device(self, device) -> torch.Tensor:
        """perm uploaded once per (re)initialisation; the reference re-uploads it every forward."""
        if self._perm_dev is None or self._perm_dev.device != device:
            self._perm_dev = self.perm.to(device=device, dtype=torch.int64)
        return self._perm_dev

    def _workspace_for(self, x: torch.Tensor) -> torch.Tensor:
        layout = F.layout_of(x)
        key = (x.device, tuple(x.shape), x.dtype, layout)
        if self._workspace_key != key:
            n, c, h, w = x.shape
            self._workspace = F.new_workspace(n, c, h, w, F.dtype_code(x), x.device, layout)
            self._workspace_key = key
        return self._workspace

    def is_active(self) -> bool:
        return bool(self.rand_p < self.p) and not (not self.mix_style and self.no_noise)

    # ------------------------------------------------------------------------------------
    def forward(self, x):
        self.data = x
        n, c = x.size(0), x.size(1)
        plane = x.numel() // (n * c) if n * c else 0     # x.view(B, C, -1).size(2) without touching memory
        # identity cases of the reference (maxstyle.py:146-152): the SAME tensor object comes back
        if (self.rand_p >= self.p) or (not self.mix_style and self.no_noise) or n <= 1 or plane == 1:
            return x
        assert self.batch_size == n and self.num_feature == c, \
            f"check input dim, expect ({self.batch_size}, {self.num_feature}, *,*) , got {n}{c}"
        if x.dim() != 4:
            raise RuntimeError(f"maxstyle_b200: expected a 4-d [N,C,H,W] feature map, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("maxstyle_b200: MaxStyle.forward got a CPU tensor; the layer runs only as CUDA kernels "
                               "on a B200 and has no CPU fallback")
        if self.gamma_noise.device != x.device:
            raise RuntimeError(f"maxstyle_b200: parameters are on {self.gamma_noise.device}, input on {x.device}")
        with F.device_guard(x.device):
            return F.MaxStyleFunction.apply(x, self.gamma_noise, self.beta_noise, self.lmda, self, L.PRE_NONE, 0.0, None)

    # ------------------------------------------------------------------------------------
    # neighbour fusion -- an extension, not part of the reference surface
    # ------------------------------------------------------------------------------------
    def forward_fused(self, z, pre: str = "leaky_relu", negative_slope: float = 0.2, collect_minmax: bool = False):
        """`self(act(z))` without materialising act(z): the kernels apply the activation as they load z (and the backward
        returns dz).  `pre`: "leaky_relu" (the LeakyReLU(0.2) that ends the reference's res_up_family blocks,
        src/models/ebm/encoder_decoder.py:337-357) or "sigmoid" (the decoder's last_act in front of layer 5, :624-630).
        `collect_minmax=True` also returns the per-plane (min, max) of y, gathered while y is written, as the int32 tensors
        `maxstyle_b200.fused.rescale_intensity` takes (the reference rescales the loop's output, model:868-869).
        Identity cases (inactive draw, batch of 1, ...) return act(z) computed by PyTorch, like `forward` returns x."""
        if pre not in ("leaky_relu", "sigmoid"):
            raise ValueError(f"pre must be 'leaky_relu' or 'sigmoid', got {pre!r}")
        n, c = z.size(0), z.size(1)
        plane = z.numel() // (n * c) if n * c else 0
        if (self.rand_p >= self.p) or (not self.mix_style and self.no_noise) or n <= 1 or plane == 1:
            x = torch.nn.functional.leaky_relu(z, negative_slope) if pre == "leaky_relu" else torch.sigmoid(z)
            self.data = x
            return (x, None) if collect_minmax else x
        assert self.batch_size == n and self.num_feature == c, \
            f"check input dim, expect ({self.batch_size}, {self.num_feature}, *,*) , got {n}{c}"
        if not z.is_cuda:
            raise RuntimeError("maxstyle_b200: MaxStyle.forward_fused got a CPU tensor; the layer runs only as CUDA kernels "
                               "on a B200 and has no CPU fallback")
        if not z.is_contiguous():
            z = z.contiguous()                     # the fused activation is built for NCHW
        self.data = z
        minmax = None
        if collect_minmax:
            minmax = (torch.full((n * c,), -1, dtype=torch.int32, device=z.device), torch.zeros(n * c, dtype=torch.int32, device=z.device))
        op = L.PRE_LEAKY_RELU if pre == "leaky_relu" else L.PRE_SIGMOID
        with F.device_guard(z.device):
            y = F.MaxStyleFunction.apply(z, self.gamma_noise, self.beta_noise, self.lmda, self, op, float(negative_slope), minmax)
        return (y, minmax) if collect_minmax else y
