"""Builds and binds the C restatement of the layer (oracle/maxstyle_oracle.c) -- TEST INFRASTRUCTURE ONLY.

`build()` compiles oracle/_build/libmaxstyle_oracle.so with gcc (the directory is git-ignored; the .so travels to the GPU box
like the product library); `forward` / `backward` wrap it for numpy arrays.  Used by tests/test_oracle_golden.py and by
__graft_entry__.build() ("building the checker is not using it").
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "maxstyle_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libmaxstyle_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build the C oracle")
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run([gcc, "-O2", "-std=c99", "-fPIC", "-shared", "-o", LIB + ".tmp", SRC, "-lm"], check=True, capture_output=True)
    os.replace(LIB + ".tmp", LIB)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def flags_of(mix_style=True, no_noise=False, compute_std=True, no_clamp=False) -> int:
    return (1 if mix_style else 0) | (2 if no_noise else 0) | (4 if compute_std else 0) | (8 if no_clamp else 0)


def forward(x, perm, lmda, gamma_noise, beta_noise, eps=1e-6, flags=5, gamma_std=None, beta_std=None):
    """Returns (y, cache) with cache = dict(mu, sig, A, gamma_std, beta_std) as [N, C] / [C] float64 arrays."""
    n, c = x.shape[0], x.shape[1]
    m = int(np.prod(x.shape[2:]))
    xd = _d(x).reshape(n, c, m)
    y = np.empty_like(xd)
    mu, sig, a = (np.empty((n, c)) for _ in range(3))
    gs = np.zeros(c) if gamma_std is None else _d(gamma_std).reshape(c).copy()
    bs = np.zeros(c) if beta_std is None else _d(beta_std).reshape(c).copy()
    pm = np.ascontiguousarray(perm, dtype=np.int64)
    lm, gn, bn = _d(lmda).reshape(n), _d(gamma_noise).reshape(n, c), _d(beta_noise).reshape(n, c)
    rc = lib().ms_oracle_forward(_p(xd), n, c, C.c_int64(m), C.c_double(eps), _p(pm), _p(lm), _p(gn), _p(bn), _p(gs), _p(bs),
                                 flags, _p(y), _p(mu), _p(sig), _p(a))
    if rc != 0:
        raise ValueError("ms_oracle_forward: bad sizes (identity cases are handled upstream)")
    return y.reshape(x.shape), dict(mu=mu, sig=sig, A=a, gamma_std=gs, beta_std=bs, perm=pm, lmda=lm, flags=flags)


def backward(dy, x, cache):
    """Returns (dx, d_gamma [N,C], d_beta [N,C], d_lmda [N])."""
    n, c = x.shape[0], x.shape[1]
    m = int(np.prod(x.shape[2:]))
    xd, gd = _d(x).reshape(n, c, m), _d(dy).reshape(n, c, m)
    dx = np.empty_like(xd)
    dg, db, dl = np.empty((n, c)), np.empty((n, c)), np.empty(n)
    rc = lib().ms_oracle_backward(_p(gd), _p(xd), n, c, C.c_int64(m), _p(cache["perm"]), _p(cache["lmda"]), _p(cache["gamma_std"]),
                                  _p(cache["beta_std"]), cache["flags"], _p(cache["mu"]), _p(cache["sig"]), _p(cache["A"]), _p(dx),
                                  _p(dg), _p(db), _p(dl))
    if rc != 0:
        raise ValueError("ms_oracle_backward: bad sizes")
    return dx.reshape(x.shape), dg, db, dl
