#!/usr/bin/env python
"""Golden vectors for the pixel-wise cross entropy from the UNMODIFIED reference (src/models/custom_loss.py:1043).

Build container only:   python oracle/gen_golden_ce2d.py      # rewrites tests/golden/ce2d.npz + CE2D_MANIFEST.json
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(HERE, ".."))
try:                                   # stand-ins for modules the reference imports but the loss never touches
    from oracle import ref_shims
    ref_shims.install()
except ImportError:
    pass

CASES = [
    # (shape N,C,H,W, seed, class weights, masked, size_average, some labels = -100, logit scale)
    ((3, 4, 9, 11), 0, None, False, True, False, 1.0),
    ((2, 4, 16, 16), 1, [1.0, 2.0, 0.5, 4.0], False, True, False, 3.0),
    ((2, 3, 8, 10), 2, None, True, True, False, 1.0),
    ((2, 5, 7, 7), 3, [0.2, 0.3, 0.1, 0.2, 0.2], True, False, False, 10.0),
    ((4, 2, 6, 6), 4, None, False, True, True, 1.0),
    ((1, 1, 5, 5), 5, None, False, True, False, 1.0),          # one class: loss 0
    ((2, 4, 12, 12), 6, None, False, False, False, 30.0),       # large logits: log-sum-exp stability
]


SOFT_CASES = [
    # soft-target branch (custom_loss.py:1079-1102): (shape, seed, class weights, masked, size_average, is_gt, logit scale)
    ((3, 4, 9, 11), 0, None, False, True, False, 1.0),
    ((2, 4, 16, 16), 1, [1.0, 2.0, 0.5, 4.0], False, True, False, 3.0),
    ((2, 3, 8, 10), 2, None, True, True, True, 1.0),
    ((2, 5, 7, 7), 3, [0.2, 0.3, 0.1, 0.2, 0.2], True, False, True, 10.0),
    ((2, 8, 6, 6), 4, None, False, False, False, 30.0),
]


def soft_main():
    from src.models.custom_loss import cross_entropy_2D
    arrays, manifest = {}, []
    for idx, (shape, seed, weights, masked, size_average, is_gt, scale) in enumerate(SOFT_CASES):
        n, c, h, w = shape
        rs = np.random.RandomState(1900 + seed)
        logits = (rs.standard_normal(size=shape) * scale).astype(np.float32)
        target = (rs.standard_normal(size=shape) * scale).astype(np.float32)
        if is_gt:                                                   # probabilities over the class axis
            e = np.exp(target - target.max(axis=1, keepdims=True))
            target = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
        mask = (rs.uniform(size=(n, 1, h, w)) < 0.7).astype(np.float32) if masked else None
        x = torch.from_numpy(logits.copy()).requires_grad_(True)
        t = torch.from_numpy(target.copy()).requires_grad_(True)
        loss = cross_entropy_2D(x, t, weight=weights, size_average=size_average,
                                mask=None if mask is None else torch.from_numpy(mask.copy()), is_gt=is_gt)
        (loss * 1.5).backward()
        rec = dict(logits=logits, target=target, loss=np.float32(loss.item()), dlogits=x.grad.numpy().copy(), dtarget=t.grad.numpy().copy())
        if mask is not None:
            rec["mask"] = mask
        for k, v in rec.items():
            arrays[f"c{idx}_{k}"] = v
        manifest.append(dict(idx=idx, shape=list(shape), weights=weights, masked=masked, size_average=size_average, is_gt=is_gt,
                             scale=scale, dloss=1.5))
    np.savez_compressed(os.path.join(OUT_DIR, "ce2d_soft.npz"), **arrays)
    with open(os.path.join(OUT_DIR, "CE2D_SOFT_MANIFEST.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, reference="src/models/custom_loss.py:cross_entropy_2D (4-d target)", cases=manifest), f, indent=1)
    print("wrote", len(SOFT_CASES), "soft-target cases")


def main():
    from src.models.custom_loss import cross_entropy_2D
    arrays, manifest = {}, []
    for idx, (shape, seed, weights, masked, size_average, ignore, scale) in enumerate(CASES):
        n, c, h, w = shape
        rs = np.random.RandomState(900 + seed)
        logits = (rs.standard_normal(size=shape) * scale).astype(np.float32)
        target = rs.randint(0, c, size=(n, h, w)).astype(np.int64)
        if ignore:
            target[rs.uniform(size=target.shape) < 0.2] = -100
        mask = (rs.uniform(size=(n, 1, h, w)) < 0.7).astype(np.float32) if masked else None
        x = torch.from_numpy(logits.copy()).requires_grad_(True)
        loss = cross_entropy_2D(x, torch.from_numpy(target), weight=weights, size_average=size_average,
                                mask=None if mask is None else torch.from_numpy(mask.copy()))
        (loss * 1.5).backward()                                    # upstream gradient 1.5
        rec = dict(logits=logits, target=target, loss=np.float32(loss.item()), dlogits=x.grad.numpy().copy())
        if mask is not None:
            rec["mask"] = mask
        for k, v in rec.items():
            arrays[f"c{idx}_{k}"] = v
        manifest.append(dict(idx=idx, shape=list(shape), weights=weights, masked=masked, size_average=size_average, ignore=ignore,
                             scale=scale, dloss=1.5))
    np.savez_compressed(os.path.join(OUT_DIR, "ce2d.npz"), **arrays)
    with open(os.path.join(OUT_DIR, "CE2D_MANIFEST.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, reference="src/models/custom_loss.py:cross_entropy_2D", cases=manifest), f, indent=1)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    if "--soft-only" not in sys.argv:
        main()
    soft_main()
