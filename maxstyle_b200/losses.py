"""Pixel-wise cross entropy on NCHW logits -- drop-in for the reference's `cross_entropy_2D`
(src/models/custom_loss.py:1043-1105; called through `basic_loss_fn(..., loss_type='cross entropy')` at :23-26, which is the
loss whose NEGATIVE the inner style-optimisation loop back-propagates, advanced_triplet_recon_segmentation_model.py:555).
SURVEY.md section 8f-4.  Same name, arguments and result (a 0-dim tensor); one kernel forward, one backward
(`maxstyle_ce2d_fwd/bwd`) instead of log_softmax + two transposes + a contiguous copy + nll_loss + mask + sum.

Both branches: a label map (3-d int64 target, :1069-1078) and a soft target (4-d, logits or -- `is_gt` -- probabilities,
:1079-1102), with optional class weights (normalised to sum C like the reference), optional mask, size_average.  The loss and
its gradients come out of ONE kernel (`maxstyle_ce2d_fwd_grad`); backward only scales them by the upstream gradient.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import functional as F


_WORKSPACES = {}          # device -> zero-filled scratch of the loss kernels (they leave it zeroed; calls on one stream are ordered)
_MAX_FUSED_CLASSES = 8    # kCeMaxC: classes the one-kernel loss + gradient keeps in registers


def _workspace(device, nbytes: int) -> torch.Tensor:
    ws = _WORKSPACES.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _WORKSPACES[device] = ws
    return ws


class _CrossEntropy2D(torch.autograd.Function):
    """loss = cross_entropy_2D(logits, target).  `kind`: 0 label map (int64 [N,H,W]), 1 soft target given as logits,
    2 soft target given as probabilities (is_gt).  ONE kernel computes the loss and, when something requires grad, the
    gradients for an upstream gradient of 1 (maxstyle_ce2d_fwd_grad); backward scales them in place by the real upstream gradient
    (maxstyle_ce2d_scale, which reads it on the device: no host synchronisation)."""

    @staticmethod
    def forward(ctx, logits, target, weight, mask, size_average, kind):
        n, c, h, w = logits.shape
        lib = L.get_lib()
        need_dl = logits.requires_grad
        need_dt = kind != 0 and target.requires_grad
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits) if need_dl else None
        dtarget = torch.empty_like(target) if need_dt else None
        if c <= _MAX_FUSED_CLASSES:
            ws = _workspace(logits.device, int(lib.maxstyle_ce2d_workspace_bytes(n, c, h, w)))
            rc = lib.maxstyle_ce2d_fwd_grad(logits.data_ptr(), target.data_ptr() if kind == 0 else None,
                                            target.data_ptr() if kind != 0 else None, int(kind == 2), F._ptr(weight), F._ptr(mask),
                                            loss.data_ptr(), F._ptr(dlogits), F._ptr(dtarget), n, c, h, w, F.dtype_code(logits),
                                            int(bool(size_average)), ws.data_ptr(), ws.numel(), F._stream())
            L.check(rc, "maxstyle_ce2d_fwd_grad")
            F.launches.kernels += 1
            ctx.fused = True
            ctx.grads = (dlogits, dtarget)
        else:
            if kind != 0:
                raise NotImplementedError(f"maxstyle_b200: cross_entropy_2D with a soft target supports up to {_MAX_FUSED_CLASSES} classes, got {c}")
            ws = _workspace(logits.device, int(lib.maxstyle_ce2d_workspace_bytes(n, c, h, w)))
            rc = lib.maxstyle_ce2d_fwd(logits.data_ptr(), target.data_ptr(), F._ptr(weight), F._ptr(mask), loss.data_ptr(), n, c, h, w,
                                       F.dtype_code(logits), int(bool(size_average)), ws.data_ptr(), ws.numel(), F._stream())
            L.check(rc, "maxstyle_ce2d_fwd")
            F.launches.kernels += 1
            ctx.fused = False
            ctx.save_for_backward(logits, target)
        ctx.aux = (weight, mask, bool(size_average))
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dloss):
        lib = L.get_lib()
        g = dloss.to(dtype=torch.float32).contiguous()
        if ctx.fused:
            out = []
            for t in ctx.grads:
                if t is not None:
                    with F.device_guard(t.device):
                        rc = lib.maxstyle_ce2d_scale(t.data_ptr(), g.data_ptr(), t.numel(), F.dtype_code(t), F._stream())
                    L.check(rc, "maxstyle_ce2d_scale")
                    F.launches.kernels += 1
                out.append(t)
            ctx.grads = (None, None)               # the buffers now belong to autograd
            return out[0], out[1], None, None, None, None
        logits, target = ctx.saved_tensors
        weight, mask, size_average = ctx.aux
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None, None
        n, c, h, w = logits.shape
        dlogits = torch.empty_like(logits)
        with F.device_guard(logits.device):
            rc = lib.maxstyle_ce2d_bwd(logits.data_ptr(), target.data_ptr(), F._ptr(weight), F._ptr(mask), g.data_ptr(),
                                       dlogits.data_ptr(), n, c, h, w, F.dtype_code(logits), int(size_average), F._stream())
        L.check(rc, "maxstyle_ce2d_bwd")
        F.launches.kernels += 1
        return dlogits, None, None, None, None, None


def cross_entropy_2D(input, target, weight=None, size_average=True, mask=None, is_gt=False):
    """Cross entropy of 4-d NCHW logits against a 3-d label map (custom_loss.py:1069-1078) or a 4-d soft target
    (custom_loss.py:1079-1102: logits, or probabilities when `is_gt`), averaged over N*H*W when `size_average`.

    weight: per-class weights (any sequence / array / tensor of length C), normalised to sum C as in the reference (:1066-1068).
    mask: [N,1,H,W] (or anything with N*H*W elements); entries with 0 are skipped; the divisor stays N*H*W (:1064, :1077).
    A label outside [0, C) other than -100 (ignored) stops the kernel with a CUDA device-side trap, like the reference's nll_loss."""
    if input.dim() != 4:
        raise RuntimeError(f"maxstyle_b200: cross_entropy_2D expects 4-d NCHW logits, got {tuple(input.shape)}")
    if target.dim() not in (3, 4):
        raise NotImplementedError
    n, c, h, w = input.shape
    logits = input.contiguous()
    if target.dim() == 3:
        if target.numel() != n * h * w:
            raise RuntimeError(f"maxstyle_b200: target has {target.numel()} labels, logits have {n * h * w} pixels")
        tgt = target.to(device=input.device, dtype=torch.int64).contiguous()
        kind = 0
    else:
        if tuple(target.shape) != tuple(input.shape):
            raise RuntimeError(f"maxstyle_b200: soft target {tuple(target.shape)} does not match the logits {tuple(input.shape)}")
        tgt = target.to(device=input.device, dtype=input.dtype).contiguous()
        kind = 2 if is_gt else 1
    if not input.is_cuda:
        raise RuntimeError("maxstyle_b200: cross_entropy_2D got a CPU tensor; it runs only as CUDA kernels on a B200 "
                           "(there is no CPU path)")
    wt = None
    if weight is not None:
        wnp = np.array(weight.detach().cpu() if isinstance(weight, torch.Tensor) else weight, dtype=np.float64)
        wnp = wnp / (1.0 * wnp.sum()) * c
        wt = torch.tensor(wnp, device=input.device, dtype=torch.float32)
        if wt.numel() != c:
            raise RuntimeError(f"maxstyle_b200: {wt.numel()} class weights for {c} classes")
    mk = None
    if mask is not None:
        mk = mask.detach().to(device=input.device, dtype=torch.float32).reshape(-1).contiguous()
        if mk.numel() != n * h * w:
            raise RuntimeError(f"maxstyle_b200: mask has {mk.numel()} entries, logits have {n * h * w} pixels")
    with F.device_guard(input.device):
        return _CrossEntropy2D.apply(logits, tgt, wt, mk, size_average, kind)
