"""Pixel-wise cross entropy on NCHW logits -- drop-in for the reference's `cross_entropy_2D`
(src/models/custom_loss.py:1043-1105; called through `basic_loss_fn(..., loss_type='cross entropy')` at :23-26, which is the
loss whose NEGATIVE the inner style-optimisation loop back-propagates, advanced_triplet_recon_segmentation_model.py:555).
SURVEY.md section 8f-4.  Same name, arguments and result (a 0-dim tensor); one kernel forward, one backward
(`maxstyle_ce2d_fwd/bwd`) instead of log_softmax + two transposes + a contiguous copy + nll_loss + mask + sum.

Built: the label-map branch (3-d int64 target), optional class weights (normalised to sum C like the reference), optional mask,
size_average.  Not built: the soft-target branch (4-d target, :1079-1102), which the loop does not use -- it raises
NotImplementedError here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import functional as F


class _CrossEntropy2D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weight, mask, size_average):
        n, c, h, w = logits.shape
        lib = L.get_lib()
        ws = torch.zeros(int(lib.maxstyle_ce2d_workspace_bytes(n, c, h, w)), dtype=torch.uint8, device=logits.device)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        rc = lib.maxstyle_ce2d_fwd(logits.data_ptr(), target.data_ptr(), F._ptr(weight), F._ptr(mask), loss.data_ptr(), n, c, h, w,
                                   F.dtype_code(logits), int(bool(size_average)), ws.data_ptr(), ws.numel(), F._stream())
        L.check(rc, "maxstyle_ce2d_fwd")
        F.launches.kernels += 1
        ctx.save_for_backward(logits, target)
        ctx.aux = (weight, mask, bool(size_average))
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dloss):
        logits, target = ctx.saved_tensors
        weight, mask, size_average = ctx.aux
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        n, c, h, w = logits.shape
        dlogits = torch.empty_like(logits)
        g = dloss.to(dtype=torch.float32).contiguous()
        with torch.cuda.device(logits.device):
            rc = L.get_lib().maxstyle_ce2d_bwd(logits.data_ptr(), target.data_ptr(), F._ptr(weight), F._ptr(mask), g.data_ptr(),
                                               dlogits.data_ptr(), n, c, h, w, F.dtype_code(logits), int(size_average), F._stream())
        L.check(rc, "maxstyle_ce2d_bwd")
        F.launches.kernels += 1
        return dlogits, None, None, None, None


def cross_entropy_2D(input, target, weight=None, size_average=True, mask=None, is_gt=False):
    """Cross entropy of 4-d NCHW logits against a 3-d label map, averaged over N*H*W (custom_loss.py:1043).

    weight: per-class weights (any sequence / array / tensor of length C), normalised to sum C as in the reference (:1066-1068).
    mask: [N,1,H,W] (or anything with N*H*W elements); entries with 0 are skipped; the divisor stays N*H*W (:1064, :1077)."""
    if input.dim() != 4:
        raise RuntimeError(f"maxstyle_b200: cross_entropy_2D expects 4-d NCHW logits, got {tuple(input.shape)}")
    if target.dim() == 4:
        raise NotImplementedError("maxstyle_b200: the soft-target branch of cross_entropy_2D (4-d target) is not built; "
                                  "the MaxStyle loop uses label maps")
    if target.dim() != 3:
        raise NotImplementedError
    if not input.is_cuda:
        raise RuntimeError("maxstyle_b200: cross_entropy_2D got a CPU tensor; it runs only as CUDA kernels on a B200 "
                           "(there is no CPU path)")
    n, c, h, w = input.shape
    if target.numel() != n * h * w:
        raise RuntimeError(f"maxstyle_b200: target has {target.numel()} labels, logits have {n * h * w} pixels")
    logits = input.contiguous()
    tgt = target.to(device=input.device, dtype=torch.int64).contiguous()
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
    with torch.cuda.device(input.device):
        return _CrossEntropy2D.apply(logits, tgt, wt, mk, size_average)
