"""CPU oracle for the pixel-wise cross entropy -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's ``cross_entropy_2D`` (src/models/custom_loss.py:1043-1105): the label-map branch
(:1069-1078) and the soft-target branch (:1079-1102), pinned against outputs of the unmodified reference
(``oracle/gen_golden_ce2d.py`` -> tests/golden/ce2d.npz, ce2d_soft.npz).
"""
from __future__ import annotations

import numpy as np

__all__ = ["cross_entropy_2d", "cross_entropy_2d_grad", "cross_entropy_2d_soft", "cross_entropy_2d_soft_grad"]


def _prep(logits, target, weight, mask, dtype):
    n, c, h, w = logits.shape
    l = np.asarray(logits, dtype=dtype).transpose(0, 2, 3, 1).reshape(-1, c)           # (:1060) NHWC rows
    t = np.asarray(target).reshape(-1)
    wt = np.ones(c, dtype=dtype)
    if weight is not None:
        wt = np.asarray(weight, dtype=np.float64)
        wt = (wt / (1.0 * wt.sum()) * c).astype(np.float32).astype(dtype)                # (:1066-1068), float32 tensor
    m = np.ones(n * h * w, dtype=dtype) if mask is None else np.asarray(mask, dtype=dtype).reshape(-1)   # (:1061-1063)
    mx = l.max(axis=1, keepdims=True)
    lse = np.log(np.exp(l - mx).sum(axis=1, keepdims=True)) + mx
    logp = l - lse                                                                        # (:1059) log_softmax
    return n, c, h, w, logp, t, wt, m


def cross_entropy_2d(logits, target, weight=None, size_average=True, mask=None, dtype=np.float64):
    """loss = -sum_p mask_p * w[t_p] * logp[p, t_p] / (N*H*W if size_average)   (:1069-1078); labels == -100 are ignored
    (default ignore_index of F.nll_loss, :1071)."""
    n, c, h, w, logp, t, wt, m = _prep(logits, target, weight, mask, dtype)
    valid = t != -100
    tt = np.where(valid, t, 0)
    per = -wt[tt] * logp[np.arange(logp.shape[0]), tt] * valid
    loss = (per * m).sum()
    if size_average:
        loss = loss / float(n * h * w)                                                    # mask_region_size = numel(mask) (:1064)
    return dtype(loss)


def cross_entropy_2d_grad(logits, target, weight=None, size_average=True, mask=None, dloss=1.0, dtype=np.float64):
    n, c, h, w, logp, t, wt, m = _prep(logits, target, weight, mask, dtype)
    valid = t != -100
    tt = np.where(valid, t, 0)
    sm = np.exp(logp)
    onehot = np.zeros_like(sm)
    onehot[np.arange(sm.shape[0]), tt] = 1.0
    g = (sm - onehot) * (wt[tt] * m * valid)[:, None] * dloss
    if size_average:
        g = g / float(n * h * w)
    return g.reshape(n, h, w, c).transpose(0, 3, 1, 2).astype(dtype)


def _soft_prep(logits, target, weight, mask, is_gt, dtype):
    n, c, h, w = logits.shape
    l = np.asarray(logits, dtype=dtype).transpose(0, 2, 3, 1).reshape(-1, c)
    t = np.asarray(target, dtype=dtype).transpose(0, 2, 3, 1).reshape(-1, c)           # (:1085-1086)
    wt = np.ones(c, dtype=dtype)
    if weight is not None:
        wt = np.asarray(weight, dtype=np.float64)
        wt = (wt / (1.0 * wt.sum()) * c).astype(np.float32).astype(dtype)
    m = np.ones(n * h * w, dtype=dtype) if mask is None else np.asarray(mask, dtype=dtype).reshape(-1)
    mx = l.max(axis=1, keepdims=True)
    logp = l - (np.log(np.exp(l - mx).sum(axis=1, keepdims=True)) + mx)
    if is_gt:
        q = t                                                                            # (:1083-1084)
    else:
        tm = t.max(axis=1, keepdims=True)
        q = np.exp(t - tm)
        q = q / q.sum(axis=1, keepdims=True)                                             # (:1081-1082) softmax(target)
    return n, c, h, w, logp, q, wt, m


def cross_entropy_2d_soft(logits, target, weight=None, size_average=True, mask=None, is_gt=False, dtype=np.float64):
    """loss = -sum_p mask_p sum_c w_c q_pc logp_pc / (N*H*W if size_average), q = softmax(target) or target (:1079-1102)."""
    n, c, h, w, logp, q, wt, m = _soft_prep(logits, target, weight, mask, is_gt, dtype)
    loss = -((q * logp * wt[None, :]).sum(axis=1) * m).sum()
    if size_average:
        loss = loss / float(n * h * w)
    return dtype(loss)


def cross_entropy_2d_soft_grad(logits, target, weight=None, size_average=True, mask=None, is_gt=False, dloss=1.0, dtype=np.float64):
    """(d loss / d logits, d loss / d target) of the soft-target branch, times `dloss`."""
    n, c, h, w, logp, q, wt, m = _soft_prep(logits, target, weight, mask, is_gt, dtype)
    p = np.exp(logp)
    wq = (wt[None, :] * q).sum(axis=1, keepdims=True)
    scale = (m * dloss / (float(n * h * w) if size_average else 1.0))[:, None]
    dl = -(wt[None, :] * q - p * wq) * scale
    wl = wt[None, :] * logp
    if is_gt:
        dt = -wl * scale
    else:
        dt = -q * (wl - (q * wl).sum(axis=1, keepdims=True)) * scale
    back = lambda a: a.reshape(n, h, w, c).transpose(0, 3, 1, 2).astype(dtype)
    return back(dl), back(dt)
