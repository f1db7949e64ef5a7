"""Host-buffer entry point of the layer step: feature maps and upstream gradients that live in pinned HOST
memory go through forward + backward + fused parameter step on the GPU and come back to host buffers.

This is the call the end-to-end number of bench.py times (`e2e`).  The arithmetic is the same kernels as
`MaxStyle.forward/backward`; what this module adds is the copy schedule.  One step moves 2 tensors each
way over PCIe (x, dy in; y, dX out), which costs ~70x the kernels' time, so the schedule is what matters:

    copy-in stream :  x(i) --------- dy(i) ---------  x(i+1) -------- dy(i+1) ...
    compute stream :        fwd(i)          bwd+step(i)        fwd(i+1) ...
    copy-out stream:              y(i) ---------- dX(i), params(i) ------ y(i+1) ...

The link is full duplex: the copy-in of step i+1 runs under the copy-out of step i (two device slots for
x / dy, two host slots for the results), so in steady state a step costs max(bytes in, bytes out) / link
rate instead of their sum.  `submit()` enqueues one step and returns a ticket; `wait(ticket)` blocks until
that step's results are in its host slot.  At most `depth` (2) steps are in flight.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from .layer import MaxStyle


@dataclass
class HostStepResult:
    """Pinned host buffers holding the results of one step (valid after `wait`)."""
    y: torch.Tensor          # [N,C,H,W] layer output
    dx: torch.Tensor         # [N,C,H,W] gradient w.r.t. x
    params: torch.Tensor     # [2*N*C + N] gamma_noise | beta_noise | lmda after the step


class HostStepPipeline:
    """Forward + backward + fused step of one `MaxStyle` layer on host-resident batches.

    Args:
        layer: an active MaxStyle module on the target device (attach a FusedStyleOptimizer to it first
            if the parameter step is wanted).
        shape: (N, C, H, W) of the batches.
        dtype: element type of x / dy / y / dX.
        depth: steps in flight (device input slots and host result slots).
    """

    def __init__(self, layer: MaxStyle, shape, dtype=torch.float32, depth: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("maxstyle_b200: HostStepPipeline needs a CUDA device (there is no CPU path)")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.layer, self.shape, self.dtype, self.depth = layer, tuple(shape), dtype, depth
        dev = layer.gamma_noise.device
        if dev.type != "cuda":
            raise RuntimeError("maxstyle_b200: the layer's parameters must live on a CUDA device")
        self.device = dev
        n, c, h, w = self.shape
        self.n_params = 2 * n * c + n
        with torch.cuda.device(dev):
            self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
            self.x = [torch.empty(self.shape, dtype=dtype, device=dev).requires_grad_(True) for _ in range(depth)]
            self.dy = [torch.empty(self.shape, dtype=dtype, device=dev) for _ in range(depth)]
            self.p_dev = [torch.empty(self.n_params, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.results: List[HostStepResult] = [
            HostStepResult(torch.empty(self.shape, dtype=dtype, pin_memory=True), torch.empty(self.shape, dtype=dtype, pin_memory=True),
                           torch.empty(self.n_params, dtype=torch.float32, pin_memory=True)) for _ in range(depth)]
        self.slot_free: List[Optional[torch.cuda.Event]] = [None] * depth     # compute has finished with x/dy of the slot
        self.out_done: List[Optional[torch.cuda.Event]] = [None] * depth      # results of the slot are on the host
        self.ticket = 0
        self.h2d_bytes = 2 * self.x[0].numel() * self.x[0].element_size()
        self.d2h_bytes = self.h2d_bytes + self.n_params * 4

    def submit(self, hx: torch.Tensor, hdy: torch.Tensor) -> int:
        """Enqueue one step on pinned host tensors hx (features) and hdy (upstream gradient)."""
        for t, name in ((hx, "hx"), (hdy, "hdy")):
            if t.is_cuda or tuple(t.shape) != self.shape or t.dtype != self.dtype:
                raise RuntimeError(f"maxstyle_b200: {name} must be a host tensor of shape {self.shape} and dtype {self.dtype}")
        k = self.ticket % self.depth
        if self.out_done[k] is not None:          # the slot's previous results must have been copied out (and read)
            self.out_done[k].synchronize()
        layer, x, dy, res = self.layer, self.x[k], self.dy[k], self.results[k]
        with torch.cuda.device(self.device):
            x_ready, dy_ready = torch.cuda.Event(), torch.cuda.Event()
            y_ready, bwd_done = torch.cuda.Event(), torch.cuda.Event()
            with torch.cuda.stream(self.s_in):
                if self.slot_free[k] is not None:
                    self.s_in.wait_event(self.slot_free[k])
                with torch.no_grad():
                    x.copy_(hx, non_blocking=True)
                x_ready.record(self.s_in)
                dy.copy_(hdy, non_blocking=True)
                dy_ready.record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(x_ready)
                y = layer(x)
                y_ready.record(self.s_cmp)
                self.s_cmp.wait_event(dy_ready)
                x.grad = None
                y.backward(dy)
                dx = x.grad
                torch.cat([layer.gamma_noise.detach().flatten(), layer.beta_noise.detach().flatten(),
                           layer.lmda.detach().flatten()], out=self.p_dev[k])
                bwd_done.record(self.s_cmp)
            self.slot_free[k] = bwd_done
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(y_ready)
                res.y.copy_(y.detach(), non_blocking=True)
                y.record_stream(self.s_out)
                self.s_out.wait_event(bwd_done)
                res.dx.copy_(dx, non_blocking=True)
                dx.record_stream(self.s_out)
                res.params.copy_(self.p_dev[k], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.s_out)
            self.out_done[k] = done
        self.ticket += 1
        return self.ticket - 1

    def wait(self, ticket: int) -> HostStepResult:
        """Block until the results of `ticket` are in host memory and return their buffers (reused `depth` steps later)."""
        if not (self.ticket - self.depth <= ticket < self.ticket):
            raise RuntimeError(f"maxstyle_b200: ticket {ticket} is not in flight (next ticket {self.ticket}, depth {self.depth})")
        k = ticket % self.depth
        self.out_done[k].synchronize()
        return self.results[k]

    def drain(self):
        for ev in self.out_done:
            if ev is not None:
                ev.synchronize()
