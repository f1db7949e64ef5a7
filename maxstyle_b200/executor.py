"""CUDA-graphed inner style-optimisation loop.

The reference's `generate_max_style_image` (src/models/advanced_triplet_recon_segmentation_model.py:458-571) runs,
per training iteration: construct 3 MaxStyle modules + an Adam optimiser, then n_iter+1 decoder passes with the layers
spliced in and n_iter times {encoder + segmentation decoder, loss = -CE, backward, optimiser step}.  At the sizes of
BASELINE config 2 the style layers themselves are ~2 ms of kernels but the eager loop spends several times that in
Python, allocator and launch overhead (tests/loop_config2.py).  `StyleLoopExecutor` keeps ONE set of layers whose
random state is re-drawn IN PLACE each run (`MaxStyle.reinit_`, same generator consumption as three fresh
constructions), captures the whole loop -- the caller's decoder / loss closures included -- into a CUDA graph per
activation pattern (which of the layers drew rand_p < p; at most 2^L graphs), and replays it: no allocation, no
Python between kernels, the parameter step fused into the layers' backward epilogue, no host synchronisation.

The closures must be capturable: static shapes, no host reads of device values, frozen weights (the reference freezes
them too, model:512-514).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn

from .layer import MaxStyle
from .optim import FusedStyleOptimizer


class StyleLoopExecutor:
    """Args:
        decode: `decode(code, layers) -> image`: the decoder pass with the style layers spliced in, e.g.
            `lambda code, layers: decoder.apply_max_style(code, layers, [3, 4, 5])` -- `layers` is the `nn.ModuleDict`
            keyed by `str(index)` that the reference builds (model:527).
        loss: `loss(image, *targets) -> scalar` whose gradient the style parameters DESCEND (the reference passes -CE,
            model:555).
        batch_size, channels: N and {layer index: C} of the activations the layers see (model:523 `channel_num[i]`).
        n_iter, lr, p and the MaxStyle flags: as in `generate_max_style_image`.
        step: 'adam' (reference) or 'sign'.
    """

    def __init__(self, decode: Callable, loss: Callable, batch_size: int, channels: Dict[int, int], n_iter: int = 5,
                 lr: float = 0.1, p: float = 0.5, mix_style: bool = True, no_noise: bool = False, mix_learnable: bool = True,
                 noise_learnable: bool = True, always_use_beta: bool = False, step: str = "adam", device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("maxstyle_b200: StyleLoopExecutor needs a CUDA device (there is no CPU path)")
        self.decode, self.loss, self.n_iter, self.p = decode, loss, int(n_iter), p
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        layers = {}
        for i in sorted(channels):
            m = MaxStyle(batch_size, channels[i], p=2.0, mix_style=mix_style, no_noise=no_noise, mix_learnable=mix_learnable,
                         noise_learnable=noise_learnable, always_use_beta=always_use_beta)      # p=2: storage for the active case
            m.p = p
            layers[str(i)] = m
        self.layers = nn.ModuleDict(layers)
        self.optimizer = FusedStyleOptimizer(self.layers.values(), lr=lr, mode=step)
        self._graphs: Dict[Tuple, Tuple] = {}
        self._static_in: Optional[Tuple[torch.Tensor, ...]] = None
        self.replays = 0
        self.captures = 0

    # ------------------------------------------------------------------------------------------------------
    def _loop(self, code, targets):
        """The reference loop body (model:539-568) on the executor's layers."""
        recon = self.decode(code, self.layers)
        for _ in range(self.n_iter):
            if not any(m.is_active() and self._learnable(m) for m in self.layers.values()):
                break                                                # model:530-532, 547-548: nothing to optimise
            value = self.loss(recon, *targets)
            value.backward()                                         # the step happens in the layers' backward epilogue
            recon = self.decode(code, self.layers)
        # `MaxStyle.data` (reference attribute: the last input) would keep this call's autograd graph -- and with it the
        # parameters' AccumulateGrad nodes, which remember the stream they were created on -- alive into the next call.  A
        # capture that inherits accumulators from the warm-up stream makes autograd join an uncaptured stream
        # (cudaErrorStreamCaptureIsolation), so the graph is released here.
        for m in self.layers.values():
            m.data = None
        return recon

    @staticmethod
    def _learnable(m: MaxStyle) -> bool:
        return any(isinstance(t, nn.Parameter) and t.requires_grad for t in (m.gamma_noise, m.beta_noise, m.lmda))

    def _redraw(self) -> Tuple[bool, ...]:
        return tuple(bool(m.reinit_()) for m in self.layers.values())

    def _snapshot(self):
        return [[t.detach().clone() for t in (m.gamma_noise, m.beta_noise, m.lmda)] for m in self.layers.values()]

    @torch.no_grad()
    def _restore(self, snap):
        for m, saved in zip(self.layers.values(), snap):
            for t, s in zip((m.gamma_noise, m.beta_noise, m.lmda), saved):
                t.copy_(s)
            m._redraw_batch_std = True
            st = m._fused_step
            if st is not None:
                for t in (st.gamma_m, st.gamma_v, st.beta_m, st.beta_v, st.lmda_m, st.lmda_v, st.step_dev):
                    t.zero_()

    def _capture(self, key):
        snap = self._snapshot()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                                # warm-up outside capture (allocations, cuDNN plans, workspaces)
            self._loop(self._static_in[0], self._static_in[1:])
        torch.cuda.current_stream().wait_stream(side)
        self._restore(snap)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._loop(self._static_in[0], self._static_in[1:])
        self._graphs[key] = (graph, out)
        self.captures += 1
        self._restore(snap)                                          # capture executes nothing, but keep the state explicit
        return graph, out

    def run(self, code: torch.Tensor, *targets: torch.Tensor, eager: bool = False) -> torch.Tensor:
        """One `generate_max_style_image` call: re-draw the style state, run the loop, return the augmented image
        (detached clone, like model:571).  `eager=True` runs the same loop without a graph (debugging / parity)."""
        active = self._redraw()
        if eager:
            with torch.enable_grad():
                return self._loop(code.detach(), targets).detach().clone()
        ins = (code.detach(),) + tuple(targets)
        if self._static_in is None:
            self._static_in = tuple(t.clone() for t in ins)
        else:
            if any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(self._static_in, ins)):
                raise RuntimeError("maxstyle_b200: StyleLoopExecutor inputs must keep their shapes and dtypes between runs")
            for dst, src in zip(self._static_in, ins):
                dst.copy_(src)
        with torch.enable_grad():
            entry = self._graphs.get(active) or self._capture(active)
        graph, out = entry
        graph.replay()
        self.replays += 1
        return out.detach().clone()
