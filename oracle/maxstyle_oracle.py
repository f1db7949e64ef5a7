"""CPU oracle for the MaxStyle feature-style layer -- TEST INFRASTRUCTURE ONLY.

This is a numpy restatement of the algorithm of the reference layer
(cherise215/MaxStyle, ``src/advanced/maxstyle.py``) and of the optimiser step its
caller applies (``torch.optim.Adam``, third-party, restated from its published
algorithm).  It exists so the CUDA path can be checked on a box where
``/root/reference`` is absent.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package ``maxstyle_b200`` never does (there is a test for that).

Parity pin: the reference ships no asserted golden vectors for this path (SURVEY.md
section 8c), so the oracle is pinned against outputs of the reference itself, generated
in the build container by ``oracle/gen_golden.py`` (which imports the reference by path)
and committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every one.

Shapes follow the reference: x is [N, C, H, W]; style tables mu/sig/gamma_noise/
beta_noise are kept as [N, C] (the reference's [N, C, 1, 1] squeezed), lmda as [N],
gamma_std/beta_std as [C], perm as int64 [N].

Every function cites the reference lines it restates (paths relative to the
reference checkout).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

__all__ = [
    "StyleState", "instance_stats", "batch_std", "mix_tables", "style_affine",
    "forward", "backward", "AdamState", "adam_step", "sign_step", "is_identity_case",
]


@dataclass
class StyleState:
    """The per-call random state a MaxStyle module owns (maxstyle.py:48-122)."""
    perm: np.ndarray                 # int64 [N]            (maxstyle.py:55-58)
    gamma_noise: np.ndarray          # [N, C]               (maxstyle.py:87-91 / 76-80)
    beta_noise: np.ndarray           # [N, C]
    lmda: np.ndarray                 # [N]                  (maxstyle.py:99-110)
    rand_p: float = 0.0              # (maxstyle.py:62)
    p: float = 0.5
    mix_style: bool = True
    no_noise: bool = False
    eps: float = 1e-6                # (maxstyle.py:15)
    gamma_std: Optional[np.ndarray] = None   # [C], cached after the first forward (maxstyle.py:165-168)
    beta_std: Optional[np.ndarray] = None


def is_identity_case(state: StyleState, shape) -> bool:
    """Early-out predicate of forward (maxstyle.py:146-152): the layer returns x itself."""
    n, _c = shape[0], shape[1]
    m = int(np.prod(shape[2:]))
    return (state.rand_p >= state.p) or ((not state.mix_style) and state.no_noise) or n <= 1 or m == 1


def instance_stats(x: np.ndarray, eps: float, dtype=np.float32):
    """mu = mean over HxW, sig = sqrt(unbiased var over HxW + eps)  (maxstyle.py:157-159).

    Returned as [N, C] arrays of ``dtype``.  Accumulation is done in float64 when
    dtype is float64, else numpy's pairwise float32 summation on a two-pass formula
    (mean first, then squared deviations), which is what ATen's var does up to rounding.
    """
    n, c = x.shape[0], x.shape[1]
    flat = np.asarray(x, dtype=dtype).reshape(n, c, -1)
    m = flat.shape[2]
    mu = flat.mean(axis=2, dtype=dtype)
    dev = flat - mu[:, :, None]
    var = (dev * dev).sum(axis=2, dtype=dtype) / dtype(m - 1)      # unbiased: divisor M-1
    sig = np.sqrt(var + dtype(eps))
    return mu.astype(dtype), sig.astype(dtype)


def batch_std(table: np.ndarray, dtype=np.float32) -> np.ndarray:
    """Unbiased std over the batch dimension of an [N, C] table -> [C]
    (torch.std(sig, dim=0) / torch.std(mu, dim=0), maxstyle.py:166,168)."""
    t = np.asarray(table, dtype=dtype)
    n = t.shape[0]
    mean = t.mean(axis=0, dtype=dtype)
    dev = t - mean[None, :]
    return np.sqrt((dev * dev).sum(axis=0, dtype=dtype) / dtype(n - 1)).astype(dtype)


def mix_tables(mu, sig, perm, lmda, mix_style: bool, dtype=np.float32):
    """Style mixing with the batch-permuted partner (maxstyle.py:172-179).

    Returns (mu_mix, sig_mix, clipped_lmda) with the reference's operation order
    ``sig * (1 - l) + sig2 * l``."""
    mu = np.asarray(mu, dtype=dtype)
    sig = np.asarray(sig, dtype=dtype)
    if not mix_style:
        return mu, sig, np.zeros(mu.shape[0], dtype=dtype)
    l = np.clip(np.asarray(lmda, dtype=dtype), dtype(0), dtype(1))[:, None]   # clamp (maxstyle.py:173)
    mu2, sig2 = mu[perm], sig[perm]                                           # gather rows (maxstyle.py:174)
    one = dtype(1)
    sig_mix = sig * (one - l) + sig2 * l
    mu_mix = mu * (one - l) + mu2 * l
    return mu_mix.astype(dtype), sig_mix.astype(dtype), l[:, 0]


def style_affine(mu_mix, sig_mix, state: StyleState, gamma_std, beta_std, dtype=np.float32):
    """A = sig_mix + gamma_noise*gamma_std, B = mu_mix + beta_noise*beta_std
    (maxstyle.py:181-185); with no_noise the perturbation is dropped (:182)."""
    if state.no_noise:
        return np.asarray(sig_mix, dtype=dtype), np.asarray(mu_mix, dtype=dtype)
    a = sig_mix + np.asarray(state.gamma_noise, dtype=dtype) * np.asarray(gamma_std, dtype=dtype)[None, :]
    b = mu_mix + np.asarray(state.beta_noise, dtype=dtype) * np.asarray(beta_std, dtype=dtype)[None, :]
    return a.astype(dtype), b.astype(dtype)


@dataclass
class ForwardCache:
    mu: np.ndarray
    sig: np.ndarray
    a: np.ndarray
    b: np.ndarray
    gamma_std: np.ndarray
    beta_std: np.ndarray
    identity: bool = False


def forward(x: np.ndarray, state: StyleState, dtype=np.float32, global_mu=None, global_sig=None,
            row_offset: int = 0):
    """MaxStyle.forward (maxstyle.py:140-189).  Returns (y, cache).

    ``state.gamma_std/beta_std`` are filled on the first call and reused afterwards
    exactly like the reference's lazily cached attributes (:165-168).

    Global-batch extension (BASELINE.json, no reference counterpart): when
    ``global_mu/global_sig`` ([N_global, C]) are given, x holds rows
    [row_offset, row_offset + N_local) of the global batch, ``state.perm`` indexes the
    global batch and ``state.gamma_noise/beta_noise/lmda`` hold the local rows.  This
    equals the reference run on the concatenated batch, sliced.
    """
    if is_identity_case(state, x.shape) and global_mu is None:
        return x, ForwardCache(None, None, None, None, None, None, identity=True)
    n, c = x.shape[0], x.shape[1]
    mu, sig = instance_stats(x, state.eps, dtype)
    if global_mu is None:
        g_mu, g_sig = mu, sig
    else:
        g_mu, g_sig = np.asarray(global_mu, dtype=dtype), np.asarray(global_sig, dtype=dtype)
    if state.gamma_std is None:
        state.gamma_std = batch_std(g_sig, dtype)
    if state.beta_std is None:
        state.beta_std = batch_std(g_mu, dtype)
    rows = slice(row_offset, row_offset + n)
    if state.mix_style:
        l = np.clip(np.asarray(state.lmda, dtype=dtype), dtype(0), dtype(1))[:, None]
        partner = np.asarray(state.perm)[rows]
        one = dtype(1)
        sig_mix = sig * (one - l) + g_sig[partner] * l
        mu_mix = mu * (one - l) + g_mu[partner] * l
    else:
        sig_mix, mu_mix = sig, mu
    a, b = style_affine(mu_mix, sig_mix, state, state.gamma_std, state.beta_std, dtype)
    xf = np.asarray(x, dtype=dtype).reshape(n, c, -1)
    x_normed = (xf - mu[:, :, None]) / sig[:, :, None]                      # (maxstyle.py:161)
    y = (a[:, :, None] * x_normed + b[:, :, None]).reshape(x.shape)         # (maxstyle.py:184-185)
    return y.astype(dtype), ForwardCache(mu, sig, a, b, state.gamma_std, state.beta_std)


def backward(dy: np.ndarray, x: np.ndarray, state: StyleState, cache: ForwardCache, dtype=np.float32,
             global_mu=None, global_sig=None, row_offset: int = 0):
    """Closed-form gradient of forward; autograd derives the same thing from
    maxstyle.py:157-185 because mu/sig are detached at :160.

    Returns (dx, d_gamma_noise [N,C], d_beta_noise [N,C], d_lmda [N]).
    """
    n, c = x.shape[0], x.shape[1]
    g = np.asarray(dy, dtype=dtype).reshape(n, c, -1)
    xf = np.asarray(x, dtype=dtype).reshape(n, c, -1)
    mu, sig = cache.mu, cache.sig
    x_normed = (xf - mu[:, :, None]) / sig[:, :, None]
    dx = (g * (cache.a / sig)[:, :, None]).reshape(x.shape).astype(dtype)
    d_a = (g * x_normed).sum(axis=2, dtype=dtype)
    d_b = g.sum(axis=2, dtype=dtype)
    if state.no_noise:
        d_gamma = np.zeros((n, c), dtype=dtype)
        d_beta = np.zeros((n, c), dtype=dtype)
    else:
        d_gamma = d_a * cache.gamma_std[None, :]
        d_beta = d_b * cache.beta_std[None, :]
    if state.mix_style:
        g_mu = mu if global_mu is None else np.asarray(global_mu, dtype=dtype)
        g_sig = sig if global_sig is None else np.asarray(global_sig, dtype=dtype)
        partner = np.asarray(state.perm)[row_offset:row_offset + n]
        lm = np.asarray(state.lmda, dtype=dtype)
        mask = ((lm >= 0) & (lm <= 1)).astype(dtype)          # clamp backward passes grad on the closed interval
        d_lmda = mask * (d_a * (g_sig[partner] - sig) + d_b * (g_mu[partner] - mu)).sum(axis=1, dtype=dtype)
    else:
        d_lmda = np.zeros(n, dtype=dtype)
    return dx, d_gamma.astype(dtype), d_beta.astype(dtype), d_lmda.astype(dtype)


@dataclass
class AdamState:
    """State of torch.optim.Adam for one tensor (third-party: PyTorch, the version the
    goldens were generated with is recorded in tests/golden/MANIFEST.json)."""
    exp_avg: np.ndarray
    exp_avg_sq: np.ndarray
    step: int = 0


def adam_step(param: np.ndarray, grad: np.ndarray, st: AdamState, lr: float = 0.1,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, maximize: bool = False):
    """One torch.optim.Adam step (defaults; weight_decay=0, amsgrad=False) as applied by
    the reference caller: advanced_triplet_recon_segmentation_model.py:537 (construction)
    and :562 (step); also maxstyle.py:231-238.  Arithmetic order follows torch's
    single-tensor implementation: lerp, addcmul, bias corrections, addcdiv."""
    f = np.float32
    g = np.asarray(grad, dtype=f)
    if maximize:
        g = -g
    st.step += 1
    st.exp_avg = (st.exp_avg + (g - st.exp_avg) * f(1 - beta1)).astype(f)
    st.exp_avg_sq = (st.exp_avg_sq * f(beta2) + f(1 - beta2) * g * g).astype(f)
    bc1 = 1.0 - beta1 ** st.step
    bc2 = 1.0 - beta2 ** st.step
    step_size = lr / bc1
    denom = (np.sqrt(st.exp_avg_sq) / f(np.sqrt(bc2)) + f(eps)).astype(f)
    return (param - f(step_size) * (st.exp_avg / denom)).astype(f)


def sign_step(param: np.ndarray, grad: np.ndarray, lr: float = 0.1, ascent: bool = True):
    """Sign-gradient step named by BASELINE.json north_star item (4); the reference itself
    uses Adam on -CE (whose first step is lr*g/(|g|+1e-8), i.e. nearly this).  sign(0) = 0."""
    f = np.float32
    s = np.sign(np.asarray(grad, dtype=f))
    return (param + f(lr) * s if ascent else param - f(lr) * s).astype(f)
