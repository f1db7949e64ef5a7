"""CUDA-graphed layer step: forward, backward and the fused parameter step of ONE style layer on static buffers.

The eager module path (`MaxStyle.forward` + autograd) costs ~190-330 us of Python, autograd and allocator work per step
on the host (profiles/r01_kernel_bench.txt) against ~245 us of kernels on the config-1 shape: on a slow host the GPU starves.
`GraphedLayerStep` captures the same C-ABI calls the module makes -- `maxstyle_fwd` (or, for `GlobalBatchMaxStyle`,
statistics -> all-gather -> tables -> apply) and `maxstyle_bwd` with the step in its epilogue -- into two CUDA graphs over
caller-visible static tensors and replays them; a step then costs two graph launches on the host.  Two graphs rather than
one so that a caller (bench.py) can put events between forward and backward.

This is the steady state of the reference's inner loop (maxstyle.py:165-168: gamma_std / beta_std are computed by the first
forward of a module and cached): construction runs that first forward eagerly, the graphs hold the cached-std forward.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib as L
from . import functional as F
from .layer import MaxStyle


class GraphedLayerStep:
    """Args:
        layer: an ACTIVE MaxStyle / GlobalBatchMaxStyle (attach a FusedStyleOptimizer first if the step is wanted).
        x, dy: static input tensors (NCHW or channels_last; fp32 / bf16) -- write new data into them between replays.
        need_dx: also produce the gradient w.r.t. x (False for the first spliced layer, whose input is detached).
        exchange (GlobalBatchMaxStyle only): how the (mu | sig) rows travel between ranks -- "p2p": the fused
            exchange + tables kernel over NVLink peer memory (maxstyle_tables_p2p; collective construction); "nccl":
            all_gather_into_tensor then the table kernel; "auto": p2p when symmetric memory can be set up on every rank,
            else nccl (`self.exchange` says which).
        one_kernel (with the p2p exchange): run the whole forward as ONE kernel that also exchanges the (mu | sig) rows over peer
            memory (maxstyle_fwd_p2p: the paired forward pushes a plane's statistics into every peer's inbox the moment it has
            them) when the shape qualifies; False: statistics -> exchange + tables -> apply.  Default: on.
    Attributes: `y`, `dx` (static outputs), `grads` = (d_gamma, d_beta, d_lmda) when the layer has no fused step or it
    keeps gradients, `kernels_per_step` (for launch accounting).
    """

    def __init__(self, layer: MaxStyle, x: torch.Tensor, dy: torch.Tensor, need_dx: bool = True, exchange: str = "auto",
                 one_kernel: Optional[bool] = None):
        if not x.is_cuda:
            raise RuntimeError("maxstyle_b200: GraphedLayerStep needs CUDA tensors (there is no CPU path)")
        if not layer.is_active():
            raise RuntimeError("maxstyle_b200: the layer is inactive for this draw (rand_p >= p): nothing to capture")
        self.layer = layer
        self.x = F.dense_layout(x.detach())
        self.dy = F._match_layout(dy.detach().to(self.x.dtype), self.x)
        if self.x.data_ptr() != x.data_ptr() or self.dy.data_ptr() != dy.data_ptr():
            raise RuntimeError("maxstyle_b200: x and dy must be dense (NCHW or channels_last) with matching dtype and layout")
        n, c, h, w = self.x.shape
        dev = self.x.device
        self.distributed = hasattr(layer, "_exchange")
        self.row_offset = layer.row_offset if self.distributed else 0
        self.flags = layer._flags()
        self.ws = layer._workspace_for(self.x)
        self.perm = layer._perm_device(dev)
        fused = layer._fused_step
        self.keep = fused is None or fused.keep_grads
        self.step = fused.struct(layer.gamma_noise, layer.beta_noise, layer.lmda) if fused is not None else None
        self.y = F._like(self.x)
        self.dx = F._like(self.x) if need_dx else None
        self.grads = None
        if self.keep:
            self.grads = (torch.empty(n, c, device=dev), torch.empty(n, c, device=dev), torch.empty(n, device=dev))
        self.exchange = None
        self.peer = None
        # maxstyle_rank_barrier before the one-kernel forward: measured no gain at 2 ranks (271.6 vs 268.4 us/step,
        # profiles/r01_multi.txt) -- the forward's multi-GPU overhead is not launch skew -- so it is off unless asked for
        self.start_barrier = os.environ.get("MAXSTYLE_START_BARRIER", "0") == "1"
        if one_kernel is None and os.environ.get("MAXSTYLE_ONE_KERNEL") in ("0", "1"):
            one_kernel = os.environ["MAXSTYLE_ONE_KERNEL"] == "1"      # experiments
        if one_kernel is None:                                   # the library declines (and the three-call path runs) where it does not pay
            one_kernel = self.distributed
        self.one_kernel = None if one_kernel else False          # None: ask maxstyle_fwd_p2p on the first call
        if self.distributed:
            self.table = layer._exchange.allocate(n, c, dev)
            self.mu_all, self.sig_all = layer._exchange.views(self.table)
            self.scale = torch.empty(n, c, device=dev)
            self.shift = torch.empty(n, c, device=dev)
            self.exchange = self._setup_exchange(exchange, n, c, dev)
        else:
            self.tables = torch.empty(4, n, c, dtype=torch.float32, device=dev)
            self.mu_all, self.sig_all, self.scale, self.shift = self.tables[0], self.tables[1], self.tables[2], self.tables[3]
        with torch.no_grad(), torch.cuda.device(dev):
            if layer.gamma_std is None or layer.beta_std is None:      # the module's first forward fills the cache
                layer.gamma_std = torch.empty(1, c, 1, 1, dtype=torch.float32, device=dev)
                layer.beta_std = torch.empty(1, c, 1, 1, dtype=torch.float32, device=dev)
                self._forward(self.flags | L.FLAG_COMPUTE_BATCH_STD)
            k0 = F.launches.kernels
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                               # warm-up (communicator, function attributes)
                self._forward(self.flags)
            torch.cuda.current_stream().wait_stream(side)
            self.fwd_kernels = F.launches.kernels - k0
            self.fwd_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.fwd_graph):
                self._forward(self.flags)
            self.bwd_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.bwd_graph):
                self._backward()
        F.launches.kernels = k0 + self.fwd_kernels                      # captures launch nothing; the warm-up did run
        self.kernels_per_step = self.fwd_kernels + 1

    def _setup_exchange(self, want: str, n: int, c: int, dev) -> str:
        """Collective: every rank tries to set up the peer-memory exchange, then all agree (MIN over ranks) on using it."""
        import torch.distributed as dist
        from .distributed import PeerTableExchange
        if want not in ("auto", "p2p", "nccl"):
            raise ValueError(f"exchange must be 'auto', 'p2p' or 'nccl', got {want!r}")
        if want == "nccl":
            return "nccl"
        group = self.layer._exchange.group
        ok, err = 1, None
        try:
            self.peer = PeerTableExchange(n, c, dev, group)
        except Exception as e:                                   # noqa: BLE001  (no symmetric memory on this system)
            ok, err = 0, e
        agree = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)
        if int(agree.item()) == 1:
            return "p2p"
        self.peer = None
        if want == "p2p":
            raise RuntimeError(f"maxstyle_b200: peer-memory exchange requested but not available on every rank: {err!r}")
        return "nccl"

    # the same call sequences as MaxStyleFunction / GlobalBatchFunction, on the static buffers
    def _forward(self, flags: int):
        layer = self.layer
        if self.distributed:
            n = self.x.shape[0]
            if self.exchange == "p2p" and self.one_kernel is not False:
                # whole forward in one kernel, the exchange inside its channel finaliser (maxstyle_fwd_p2p); the ranks are
                # lined up first so that the launch skew between them is not waited out inside the L2 window
                if self.one_kernel and self.start_barrier:
                    F.rank_barrier(self.peer)
                ok = F.forward_p2p(self.peer, self.x, self.mu_all, self.sig_all, self.row_offset, self.perm, layer.lmda,
                                   layer.gamma_noise, layer.beta_noise, layer.gamma_std, layer.beta_std, flags, layer.eps, self.ws,
                                   self.scale, self.shift, self.y)
                if self.one_kernel is None and not (flags & L.FLAG_COMPUTE_BATCH_STD):
                    self.one_kernel = ok                         # decided by the first steady-state call (the first forward of a
                                                                 # layer may be declined where the steady state is not); the
                                                                 # answer depends on shapes only: the same on every rank
                if ok:
                    return
            F.instance_stats(self.x, layer.eps, self.ws, self.mu_all, self.sig_all, self.row_offset)
            if self.exchange == "p2p":                           # exchange + tables: one kernel over NVLink peer memory
                F.style_tables_p2p(self.peer, self.mu_all, self.sig_all, self.row_offset, n, self.perm, layer.lmda,
                                   layer.gamma_noise, layer.beta_noise, layer.gamma_std, layer.beta_std, flags, self.scale, self.shift)
            else:
                layer._exchange.gather(self.table, n)
                F.style_tables(self.mu_all, self.sig_all, self.row_offset, n, self.perm, layer.lmda, layer.gamma_noise,
                               layer.beta_noise, layer.gamma_std, layer.beta_std, flags, self.scale, self.shift)
            F.style_apply(self.x, self.mu_all, self.row_offset, self.scale, self.shift, out=self.y)
        else:
            F.forward_raw(self.x, self.perm, layer.lmda, layer.gamma_noise, layer.beta_noise, layer.gamma_std, layer.beta_std,
                          flags, layer.eps, self.ws, out=self.y, tables=self.tables)

    def _backward(self):
        layer = self.layer
        F.backward_raw(self.dy, self.x, self.mu_all, self.sig_all, self.row_offset, self.scale, self.perm, layer.lmda,
                       layer.gamma_std, layer.beta_std, self.flags, self.ws, need_dx=self.dx is not None,
                       need_noise_grad=self.keep, need_mix_grad=self.keep, step=self.step, dx_out=self.dx, grads_out=self.grads)

    def close(self):
        """Release the two graphs.  With a distributed layer they hold captured NCCL kernels: call this (after a device
        synchronize) BEFORE `torch.distributed.destroy_process_group()`, which otherwise waits for them."""
        self.fwd_graph = None
        self.bwd_graph = None

    def forward(self) -> torch.Tensor:
        self.fwd_graph.replay()
        F.launches.kernels += self.fwd_kernels
        return self.y

    def backward(self) -> Optional[torch.Tensor]:
        self.bwd_graph.replay()
        F.launches.kernels += 1
        return self.dx

    def run(self):
        """One step: forward, backward (+ fused parameter step).  Returns (y, dx)."""
        self.forward()
        self.backward()
        return self.y, self.dx
