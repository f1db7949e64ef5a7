"""Thin tensor-level wrappers over the C ABI plus the autograd Function of the layer.

PyTorch is plumbing here: it owns the device buffers and the stream; every computation on the
path is a kernel of libmaxstyle_b200.so.  Nothing in this module computes with torch ops.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib as L

_DTYPES = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _default_sweeps():
    """Sweep flags of the three streaming kernels (include/maxstyle_b200.h).  The apply kernel walks x
    in the opposite direction to the statistics kernel and the backward walks it forward again, so each
    kernel starts on the part of x its predecessor left in L2.  MAXSTYLE_SWEEP="stats,apply,bwd"
    (three integers) overrides them for experiments."""
    import os
    env = os.environ.get("MAXSTYLE_SWEEP")
    if env:
        a, b, c = (int(v) for v in env.split(","))
        return a, b, c
    return L.SWEEP_X_KEEP, L.SWEEP_REVERSE | L.SWEEP_X_KEEP, L.SWEEP_X_STREAM


SWEEP_STATS, SWEEP_APPLY, SWEEP_BWD = _default_sweeps()


class _LaunchCounter:
    """Counts kernels of libmaxstyle_b200.so enqueued through this module (bench.py's gpu_launches)."""
    kernels = 0


launches = _LaunchCounter()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_NULL_GUARD = contextlib.nullcontext()


def _stream():
    """Raw handle of the current stream of the current device.  torch.cuda.current_stream() builds a Stream object through
    several Python layers (~20 us per call, twice per step on the eager path); the raw accessor is ~1 us."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def device_guard(device: torch.device):
    """`with torch.cuda.device(device)` only when `device` is not already current (the context manager costs ~10 us)."""
    idx = device.index
    if idx is None or idx == torch._C._cuda_getDevice():
        return _NULL_GUARD
    return torch.cuda.device(device)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"maxstyle_b200: {what} is on {t.device}; this layer only runs as CUDA kernels on a B200 "
                           "(there is no CPU path)")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"maxstyle_b200: unsupported dtype {t.dtype} (float32 and bfloat16 are implemented)")


def layout_of(x: torch.Tensor) -> int:
    """Memory layout code of a dense 4-d feature map.  NCHW-contiguous wins when both hold (C == 1 or H*W == 1)."""
    if x.is_contiguous():
        return L.NCHW
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last):
        return L.NHWC
    raise RuntimeError("maxstyle_b200: feature maps must be NCHW-contiguous or channels_last (call dense_layout first)")


def dense_layout(x: torch.Tensor) -> torch.Tensor:
    """x itself when it is NCHW-contiguous or channels_last (the reference accepts both and keeps the format);
    anything else is made NCHW-contiguous."""
    if x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)):
        return x
    return x.contiguous()


def _like(x: torch.Tensor) -> torch.Tensor:
    """Uninitialised tensor with x's shape, dtype and memory layout."""
    fmt = torch.contiguous_format if layout_of(x) == L.NCHW else torch.channels_last
    return torch.empty_like(x, memory_format=fmt)


def _match_layout(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    fmt = torch.contiguous_format if layout_of(like) == L.NCHW else torch.channels_last
    return t.contiguous(memory_format=fmt)


def workspace_bytes(n: int, c: int, h: int, w: int, dtype: int, layout: int = L.NCHW) -> int:
    return int(L.get_lib().maxstyle_workspace_bytes(n, c, h, w, dtype, layout))


_FWD_KERNELS = {}


def fwd_kernel_count(n: int, c: int, h: int, w: int, dtype: int, layout: int = L.NCHW) -> int:
    key = (n, c, h, w, dtype, layout, SWEEP_STATS)
    if key not in _FWD_KERNELS:
        _FWD_KERNELS[key] = int(L.get_lib().maxstyle_fwd_kernels(n, c, h, w, dtype, layout, SWEEP_STATS))
    return _FWD_KERNELS[key]


def workspace_status(workspace: torch.Tensor, n: int, c: int, h: int, w: int, dtype: int, layout: int = L.NCHW) -> None:
    """Synchronises; raises if a device-side wait of the fused forward timed out (debug / tests)."""
    rc = L.get_lib().maxstyle_workspace_status(workspace.data_ptr(), workspace.numel(), n, c, h, w, dtype, layout, _stream())
    L.check(rc, "maxstyle_workspace_status")


def new_workspace(n: int, c: int, h: int, w: int, dtype: int, device, layout: int = L.NCHW) -> torch.Tensor:
    """Zero-filled scratch (the library keeps it zeroed between calls)."""
    nbytes = workspace_bytes(n, c, h, w, dtype, layout)
    if nbytes == 0:
        raise RuntimeError(f"maxstyle_b200: no kernel for shape {(n, c, h, w)} dtype {dtype} layout {layout}")
    return torch.zeros(nbytes, dtype=torch.uint8, device=device)


@dataclass
class StepConfig:
    """Optimiser step fused into the backward epilogue (maxstyle_step_t)."""
    mode: int = L.STEP_ADAM
    lr: float = 0.1
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    maximize: bool = False


class FusedStepState:
    """Adam moments + device step counter for the three parameters of one layer."""

    def __init__(self, cfg: StepConfig, gamma_noise, beta_noise, lmda, update_noise: bool, update_mix: bool):
        self.cfg = cfg
        self.update_noise, self.update_mix = bool(update_noise), bool(update_mix)
        dev = lmda.device
        z = lambda t: torch.zeros_like(t, memory_format=torch.contiguous_format)
        self.gamma_m, self.gamma_v = z(gamma_noise), z(gamma_noise)
        self.beta_m, self.beta_v = z(beta_noise), z(beta_noise)
        self.lmda_m, self.lmda_v = z(lmda), z(lmda)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.keep_grads = False

    def struct(self, gamma_noise, beta_noise, lmda) -> L.StepStruct:
        c = self.cfg
        return L.StepStruct(
            mode=c.mode, maximize=int(c.maximize), lr=c.lr, beta1=c.beta1, beta2=c.beta2, eps=c.eps, t=0,
            update_noise=int(self.update_noise), update_mix=int(self.update_mix), reserved=0,
            step_dev=self.step_dev.data_ptr(),
            gamma_noise=_ptr(gamma_noise), beta_noise=_ptr(beta_noise), lmda=_ptr(lmda),
            gamma_m=_ptr(self.gamma_m), gamma_v=_ptr(self.gamma_v), beta_m=_ptr(self.beta_m), beta_v=_ptr(self.beta_v),
            lmda_m=_ptr(self.lmda_m), lmda_v=_ptr(self.lmda_v))


# --------------------------------------------------------------------------------------------
# raw calls (used by the autograd Function, the distributed layer, tests and bench)
# --------------------------------------------------------------------------------------------
def table_ld(mu_all: torch.Tensor) -> int:
    """Floats between consecutive rows of a style table (it may be a column slice of a wider buffer)."""
    if mu_all.dim() != 2 or mu_all.stride(1) != 1:
        raise RuntimeError("maxstyle_b200: style tables must be 2-d with unit stride along channels")
    return int(mu_all.stride(0))


def instance_stats(x: torch.Tensor, eps: float, workspace: torch.Tensor, mu_all=None, sig_all=None, row_offset: int = 0,
                   sweep: Optional[int] = None):
    """Kernel 1.  Returns (mu_all, sig_all) with rows [row_offset, row_offset+N) filled."""
    _require_cuda(x, "x")
    n, c, h, w = x.shape
    if mu_all is None:
        mu_all = torch.empty(n, c, dtype=torch.float32, device=x.device)
        sig_all = torch.empty(n, c, dtype=torch.float32, device=x.device)
    rc = L.get_lib().maxstyle_stats(x.data_ptr(), mu_all.data_ptr(), sig_all.data_ptr(), table_ld(mu_all), row_offset,
                                    n, c, h, w, dtype_code(x), layout_of(x), eps, SWEEP_STATS if sweep is None else sweep,
                                    workspace.data_ptr(), workspace.numel(), _stream())
    L.check(rc, "maxstyle_stats")
    launches.kernels += 1
    return mu_all, sig_all


def style_tables(mu_all, sig_all, row_offset: int, n_local: int, perm_dev, lmda, gamma_noise, beta_noise,
                 gamma_std, beta_std, flags: int, scale=None, shift=None):
    n_global, c = mu_all.shape
    if scale is None:
        scale = torch.empty(n_local, c, dtype=torch.float32, device=mu_all.device)
        shift = torch.empty(n_local, c, dtype=torch.float32, device=mu_all.device)
    rc = L.get_lib().maxstyle_tables(mu_all.data_ptr(), sig_all.data_ptr(), table_ld(mu_all), n_global, row_offset,
                                     n_local, c, _ptr(perm_dev), _ptr(lmda), _ptr(gamma_noise), _ptr(beta_noise),
                                     _ptr(gamma_std), _ptr(beta_std), flags, scale.data_ptr(), shift.data_ptr(), _stream())
    L.check(rc, "maxstyle_tables")
    launches.kernels += 1
    return scale, shift


def style_tables_p2p(peer, mu_all, sig_all, row_offset: int, n_local: int, perm_dev, lmda, gamma_noise, beta_noise,
                     gamma_std, beta_std, flags: int, scale=None, shift=None):
    """maxstyle_tables_p2p: exchange of the (mu | sig) rows over NVLink peer memory + the table step, one kernel.
    `peer` is a distributed.PeerTableExchange (symmetric-memory buffers, device epoch counter)."""
    n_global, c = mu_all.shape
    if scale is None:
        scale = torch.empty(n_local, c, dtype=torch.float32, device=mu_all.device)
        shift = torch.empty(n_local, c, dtype=torch.float32, device=mu_all.device)
    rc = L.get_lib().maxstyle_tables_p2p(peer.peers_dev.data_ptr(), peer.rank, peer.world, peer.epoch.data_ptr(),
                                         peer.done.data_ptr(), peer.error.data_ptr(), mu_all.data_ptr(), sig_all.data_ptr(),
                                         table_ld(mu_all), n_global, row_offset, n_local, c, _ptr(perm_dev), _ptr(lmda),
                                         _ptr(gamma_noise), _ptr(beta_noise), _ptr(gamma_std), _ptr(beta_std), flags,
                                         scale.data_ptr(), shift.data_ptr(), _stream())
    L.check(rc, "maxstyle_tables_p2p")
    launches.kernels += 1
    return scale, shift


def rank_barrier(peer) -> None:
    """maxstyle_rank_barrier: line the ranks up (stream-ordered, on the device) before the one-kernel multi-GPU forward."""
    rc = L.get_lib().maxstyle_rank_barrier(peer.peers_dev.data_ptr(), peer.rank, peer.world, peer.n_local, peer.channels,
                                           peer.bar_epoch.data_ptr(), peer.error.data_ptr(), _stream())
    L.check(rc, "maxstyle_rank_barrier")
    launches.kernels += 1


def forward_p2p(peer, x, mu_all, sig_all, row_offset: int, perm_dev, lmda, gamma_noise, beta_noise, gamma_std, beta_std,
                flags: int, eps: float, workspace, scale, shift, out) -> bool:
    """maxstyle_fwd_p2p: the multi-GPU forward as ONE kernel (L2-window forward whose channel finaliser exchanges the rows
    over peer memory).  Returns False -- nothing launched -- when the shape does not qualify (same answer on every rank)."""
    n, c, h, w = x.shape
    rc = L.get_lib().maxstyle_fwd_p2p(x.data_ptr(), out.data_ptr(), mu_all.data_ptr(), sig_all.data_ptr(), table_ld(mu_all),
                                      mu_all.shape[0], row_offset, _ptr(perm_dev), _ptr(lmda), _ptr(gamma_noise), _ptr(beta_noise),
                                      _ptr(gamma_std), _ptr(beta_std), scale.data_ptr(), shift.data_ptr(), n, c, h, w,
                                      dtype_code(x), layout_of(x), flags, eps, SWEEP_STATS, peer.peers_dev.data_ptr(), peer.rank,
                                      peer.world, peer.epoch.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream())
    if rc == L.ERR_UNSUPPORTED:
        return False
    L.check(rc, "maxstyle_fwd_p2p")
    launches.kernels += 1
    return True


def style_apply(x, mu_all, row_offset: int, scale, shift, out=None, sweep: Optional[int] = None):
    n, c, h, w = x.shape
    y = _like(x) if out is None else out
    rc = L.get_lib().maxstyle_apply(x.data_ptr(), y.data_ptr(), mu_all.data_ptr(), table_ld(mu_all), row_offset,
                                    scale.data_ptr(), shift.data_ptr(), n, c, h, w, dtype_code(x), layout_of(x),
                                    SWEEP_APPLY if sweep is None else sweep, _stream())
    L.check(rc, "maxstyle_apply")
    launches.kernels += 1
    return y


def forward_raw(x, perm_dev, lmda, gamma_noise, beta_noise, gamma_std, beta_std, flags: int, eps: float, workspace,
                out=None, tables=None, pre_op: int = L.PRE_NONE, pre_param: float = 0.0, minmax=None):
    """maxstyle_fwd: whole single-GPU forward.  Returns (y, mu, sig, scale, shift).
    `pre_op` != PRE_NONE: x is the PRE-activation tensor and the activation is applied as the kernels load it; `minmax`: a
    (min, max) pair of int32 [N*C] tensors initialised to -1 (0xffffffff) / 0 that receive the ordered-integer extremes of y
    (maxstyle_fwd_act)."""
    _require_cuda(x, "x")
    n, c, h, w = x.shape
    if tables is None:
        tables = torch.empty(4, n, c, dtype=torch.float32, device=x.device)
    mu, sig, scale, shift = tables[0], tables[1], tables[2], tables[3]
    y = _like(x) if out is None else out
    layout = layout_of(x)
    if pre_op != L.PRE_NONE or minmax is not None:
        rc = L.get_lib().maxstyle_fwd_act(x.data_ptr(), y.data_ptr(), mu.data_ptr(), sig.data_ptr(), _ptr(perm_dev), _ptr(lmda),
                                          _ptr(gamma_noise), _ptr(beta_noise), _ptr(gamma_std), _ptr(beta_std),
                                          scale.data_ptr(), shift.data_ptr(), n, c, h, w, dtype_code(x), layout, flags, eps,
                                          SWEEP_STATS, SWEEP_APPLY, pre_op, pre_param,
                                          None if minmax is None else minmax[0].data_ptr(), None if minmax is None else minmax[1].data_ptr(),
                                          workspace.data_ptr(), workspace.numel(), _stream())
        L.check(rc, "maxstyle_fwd_act")
        launches.kernels += 1 if (h * w * x.element_size() >= 8192 and (h * w * x.element_size()) % 16 == 0) else 3
        return y, mu, sig, scale, shift
    rc = L.get_lib().maxstyle_fwd(x.data_ptr(), y.data_ptr(), mu.data_ptr(), sig.data_ptr(), _ptr(perm_dev), _ptr(lmda),
                                  _ptr(gamma_noise), _ptr(beta_noise), _ptr(gamma_std), _ptr(beta_std),
                                  scale.data_ptr(), shift.data_ptr(), n, c, h, w, dtype_code(x), layout, flags, eps,
                                  SWEEP_STATS, SWEEP_APPLY, workspace.data_ptr(), workspace.numel(), _stream())
    L.check(rc, "maxstyle_fwd")
    launches.kernels += fwd_kernel_count(n, c, h, w, dtype_code(x), layout)   # 1 (fused) or 3 (stats + tables + apply)
    return y, mu, sig, scale, shift


def backward_raw(dy, x, mu_all, sig_all, row_offset: int, scale, perm_dev, lmda, gamma_std, beta_std, flags: int,
                 workspace, need_dx: bool = True, need_noise_grad: bool = True, need_mix_grad: bool = True,
                 step: Optional[L.StepStruct] = None, dx_out=None, grads_out=None, pre_op: int = L.PRE_NONE, pre_param: float = 0.0):
    """maxstyle_bwd.  Returns (dx | None, d_gamma | None, d_beta | None, d_lmda | None)."""
    _require_cuda(dy, "dy")
    n, c, h, w = x.shape
    dx = None
    if need_dx:
        dx = _like(x) if dx_out is None else dx_out
    dg = db = dl = None
    if grads_out is not None:
        dg, db, dl = grads_out
    else:
        if need_noise_grad:
            dg = torch.empty(n, c, dtype=torch.float32, device=x.device)
            db = torch.empty(n, c, dtype=torch.float32, device=x.device)
        if need_mix_grad:
            dl = torch.empty(n, dtype=torch.float32, device=x.device)
    if pre_op != L.PRE_NONE:
        rc = L.get_lib().maxstyle_bwd_act(dy.data_ptr(), x.data_ptr(), _ptr(dx), mu_all.data_ptr(), sig_all.data_ptr(),
                                          table_ld(mu_all), mu_all.shape[0], row_offset, scale.data_ptr(), _ptr(perm_dev), _ptr(lmda),
                                          _ptr(gamma_std), _ptr(beta_std), flags, _ptr(dg), _ptr(db), _ptr(dl),
                                          C.byref(step) if step is not None else None,
                                          n, c, h, w, dtype_code(x), layout_of(x), SWEEP_BWD, pre_op, pre_param,
                                          workspace.data_ptr(), workspace.numel(), _stream())
        L.check(rc, "maxstyle_bwd_act")
        launches.kernels += 1
        return dx, dg, db, dl
    rc = L.get_lib().maxstyle_bwd(dy.data_ptr(), x.data_ptr(), _ptr(dx), mu_all.data_ptr(), sig_all.data_ptr(),
                                  table_ld(mu_all), mu_all.shape[0], row_offset, scale.data_ptr(), _ptr(perm_dev), _ptr(lmda),
                                  _ptr(gamma_std), _ptr(beta_std), flags, _ptr(dg), _ptr(db), _ptr(dl),
                                  C.byref(step) if step is not None else None,
                                  n, c, h, w, dtype_code(x), layout_of(x), SWEEP_BWD, workspace.data_ptr(), workspace.numel(),
                                  _stream())
    L.check(rc, "maxstyle_bwd")
    launches.kernels += 1
    return dx, dg, db, dl


# --------------------------------------------------------------------------------------------
# autograd glue
# --------------------------------------------------------------------------------------------
class MaxStyleFunction(torch.autograd.Function):
    """y = MaxStyle(x; gamma_noise, beta_noise, lmda).  `layer` supplies the non-differentiable
    state (perm, cached gamma_std/beta_std, flags, workspace, optional fused step)."""

    @staticmethod
    def forward(ctx, x, gamma_noise, beta_noise, lmda, layer, pre_op=L.PRE_NONE, pre_param=0.0, minmax=None):
        x = dense_layout(x)                # NCHW or channels_last, as it came (the output keeps the format)
        flags = layer._flags()
        first = layer.gamma_std is None or layer.beta_std is None
        if first:
            gamma_std = torch.empty(1, layer.num_feature, 1, 1, dtype=torch.float32, device=x.device)
            beta_std = torch.empty(1, layer.num_feature, 1, 1, dtype=torch.float32, device=x.device)
            flags |= L.FLAG_COMPUTE_BATCH_STD
        elif getattr(layer, "_redraw_batch_std", False):       # reinit_(): recompute into the existing buffers
            gamma_std, beta_std = layer.gamma_std, layer.beta_std
            flags |= L.FLAG_COMPUTE_BATCH_STD
        else:
            gamma_std, beta_std = layer.gamma_std, layer.beta_std
        ws = layer._workspace_for(x)
        y, mu, sig, scale, shift = forward_raw(x, layer._perm_device(x.device), lmda, gamma_noise, beta_noise,
                                               gamma_std, beta_std, flags, layer.eps, ws, pre_op=pre_op, pre_param=pre_param,
                                               minmax=minmax)
        ctx.pre = (pre_op, pre_param)
        if first:                      # cached until reset(), like the reference (maxstyle.py:165-168)
            layer.gamma_std, layer.beta_std = gamma_std, beta_std
        layer._redraw_batch_std = False
        ctx.layer = layer
        ctx.flags = flags & ~L.FLAG_COMPUTE_BATCH_STD
        ctx.tables = (mu, sig, scale, gamma_std, beta_std)
        ctx.save_for_backward(x, lmda)
        ctx.set_materialize_grads(False)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        if dy is None:
            return None, None, None, None, None, None, None, None
        x, lmda = ctx.saved_tensors
        layer = ctx.layer
        mu, sig, scale, gamma_std, beta_std = ctx.tables
        need_dx, need_g, need_b, need_l = ctx.needs_input_grad[:4]
        fused = layer._fused_step
        step = fused.struct(layer.gamma_noise, layer.beta_noise, layer.lmda) if fused is not None else None
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dy = _match_layout(dy, x)
        with device_guard(x.device):
            dx, dg, db, dl = backward_raw(dy, x, mu, sig, 0, scale, layer._perm_device(x.device), lmda, gamma_std,
                                          beta_std, ctx.flags, layer._workspace_for(x), need_dx=need_dx,
                                          need_noise_grad=(need_g or need_b) and (fused is None or fused.keep_grads),
                                          need_mix_grad=need_l and (fused is None or fused.keep_grads), step=step,
                                          pre_op=ctx.pre[0], pre_param=ctx.pre[1])
        n, c = x.shape[0], x.shape[1]
        if fused is not None and not fused.keep_grads:
            return dx, None, None, None, None, None, None, None
        return (dx,
                dg.view(n, c, 1, 1) if (need_g and dg is not None) else None,
                db.view(n, c, 1, 1) if (need_b and db is not None) else None,
                dl.view(n, 1, 1, 1) if (need_l and dl is not None) else None,
                None, None, None, None)
