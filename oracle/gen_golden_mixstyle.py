#!/usr/bin/env python
"""Golden vectors for the MixStyle / DSU layer from the UNMODIFIED reference (src/advanced/mixstyle.py).

Build container only:   python oracle/gen_golden_mixstyle.py    # rewrites tests/golden/mixstyle.npz + MIXSTYLE_MANIFEST.json

The reference draws its mixing weight inside forward and does not keep it, so every case replays the reference's
generator consumption (rand(1) -> Beta sample -> randperm / randn) under the same seed to record what was drawn;
the replay is asserted against the reference's own `perm` where it keeps one.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.gen_golden import make_input, t2n, OUT_DIR  # noqa: E402

REF_FILE = "/root/reference/src/advanced/mixstyle.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("_reference_mixstyle", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MixStyle


CASES = [
    # (mix, shape, seed, fixed lmda, coefficient sampler attribute set after construction, input kind)
    ("random", (6, 3, 9, 11), 0, None, None, "scaled"),
    ("random", (4, 5, 8, 8), 1, None, None, "offset"),
    ("random", (5, 2, 12, 10), 2, 1.5, None, "scaled"),          # extrapolation: the weight is NOT clamped
    ("random", (6, 4, 6, 6), 3, None, "gaussian", "scaled"),     # Gaussian weights, some outside [0,1]
    ("random", (6, 4, 6, 6), 4, None, "uniform", "sigmoid"),
    ("crossdomain", (6, 3, 10, 10), 5, None, None, "scaled"),
    ("crossdomain", (8, 2, 7, 9), 6, None, None, "scaled"),
    ("gaussian", (6, 3, 9, 11), 7, None, None, "scaled"),
    ("gaussian", (4, 6, 8, 8), 8, None, None, "offset"),
    ("random", (6, 3, 9, 11), 9, None, None, "scaled"),          # with a caller-supplied perm (forward(x, perm))
]


def main():
    MixStyle = load_reference()
    arrays, manifest = {}, []
    for idx, (mix, shape, seed, fixed, sampler, kind) in enumerate(CASES):
        n, c = shape[0], shape[1]
        x_np = make_input(100 + seed, shape, kind)
        dy_np = np.random.RandomState(500 + seed).standard_normal(size=shape).astype(np.float32)
        m = MixStyle(p=1.0, alpha=0.1, mix=mix, lmda=fixed)
        if sampler is not None:
            m.coeficient_sampler = sampler              # the constructor ignores its argument (mixstyle.py:33)
        given_perm = torch.tensor([1, 0, 3, 2, 5, 4]) if idx == 9 else None
        x = torch.from_numpy(x_np.copy()).requires_grad_(True)
        torch.manual_seed(seed)
        y = m(x, perm=given_perm) if given_perm is not None else m(x)
        y.backward(torch.from_numpy(dy_np))
        after = float(torch.rand(1))                     # pins how much of the CPU generator the call consumed
        # replay the generator consumption
        torch.manual_seed(seed)
        p_draw = torch.rand(1)
        if fixed is None:
            if sampler in (None, "beta"):
                lm = torch.distributions.Beta(0.1, 0.1).sample((n, 1, 1, 1))
            elif sampler == "uniform":
                lm = torch.rand(n, 1, 1, 1)
            else:
                lm = torch.randn(n, 1, 1, 1)
        else:
            lm = torch.ones(n, 1, 1, 1) * fixed
        rec = dict(y=t2n(y), dx=t2n(x.grad), lmda=t2n(lm).reshape(n), rand_p=t2n(p_draw))
        if mix == "random":
            perm = given_perm if given_perm is not None else torch.randperm(n)
            assert torch.equal(perm, m.perm)
            rec["perm"] = t2n(perm).astype(np.int64)
        elif mix == "crossdomain":
            perm = torch.arange(n - 1, -1, -1)
            pb, pa = perm.chunk(2)
            pb = pb[torch.randperm(n // 2)]
            pa = pa[torch.randperm(n // 2)]
            perm = torch.cat([pb, pa], 0)
            assert torch.equal(perm, m.perm)
            rec["perm"] = t2n(perm).astype(np.int64)
        else:
            rec["eps_mu"] = t2n(torch.randn(n, c, 1, 1)).reshape(n, c)
            rec["eps_sig"] = t2n(torch.randn(n, c, 1, 1)).reshape(n, c)
        assert float(torch.rand(1)) == after, "replay consumed the generator differently from the reference"
        for k, v in rec.items():
            arrays[f"c{idx}_{k}"] = v
        manifest.append(dict(idx=idx, mix=mix, shape=list(shape), seed=seed, fixed_lmda=fixed, sampler=sampler, kind=kind,
                             x_seed=100 + seed, dy_seed=500 + seed, given_perm=given_perm is not None, next_cpu_rand=after))
    # identity branch: p > self.p returns x itself (mixstyle.py:45-48)
    m = MixStyle(p=0.0)
    x = torch.randn(3, 2, 4, 4)
    torch.manual_seed(0)
    manifest.append(dict(idx="identity", returns_same_object=bool(m(x) is x)))
    np.savez_compressed(os.path.join(OUT_DIR, "mixstyle.npz"), **arrays)
    with open(os.path.join(OUT_DIR, "MIXSTYLE_MANIFEST.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, reference_file=REF_FILE, cases=manifest), f, indent=1, sort_keys=True)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
