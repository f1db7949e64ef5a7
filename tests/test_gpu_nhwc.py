"""Parity of the NHWC (channels_last) kernels with the oracle and with the NCHW kernels.

The reference accepts a channels_last tensor and returns one: its result is
the same function of the logical [N,C,H,W] values, so the oracle is evaluated on the logical array.
Tolerances as in test_gpu_parity.py (BASELINE.json: 1e-5 forward, 1e-4 gradients).
"""
import numpy as np
import pytest
import torch

from oracle import maxstyle_oracle as O
from oracle.gen_golden import make_input
from test_gpu_parity import FWD_RTOL, GRAD_RTOL, assert_rel, dev, make_layer, n2t, oracle_state, t2n

pytestmark = pytest.mark.gpu

NHWC_SHAPES = [
    # N, C, H, W -- every dispatch of kernels_nhwc.cuh
    (4, 64, 48, 48),      # fp32 256-bit vectors, CV = 8 (shuffle stage), samples shared by several CTAs
    (3, 16, 96, 96),      # CV = 2
    (2, 8, 64, 64),       # CV = 1: whole warp folds into one channel vector
    (5, 256, 14, 14),     # CV = 32: no shuffle stage, slot = warp
    (3, 24, 40, 40),      # CV = 3: not a power of two, 255 active threads
    (6, 12, 20, 20),      # C*4 = 48 B: 128-bit vectors, CV = 3
    (4, 5, 17, 19),       # scalar path, CV = 5
    (2, 320, 9, 9),       # CV = 40: one thread row covers 6 pixels
    (40, 32, 12, 12),     # many small samples: several whole samples per CTA
    (700, 8, 4, 4),       # more samples than CTAs
    (2, 2, 300, 300),     # scalar path (C*4 = 8 B), big samples
    (20, 1, 64, 64),      # C = 1: same memory as NCHW
]


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("shape", NHWC_SHAPES)
def test_nhwc_forward_backward_vs_oracle(shape):
    n, c, h, w = shape
    seed = 17 * n + c + h
    torch.manual_seed(seed)
    layer = make_layer(n, c)
    x_np = make_input(seed, shape)
    dy_np = np.random.RandomState(seed + 1).standard_normal(size=shape).astype(np.float32)
    st = oracle_state(layer.perm.numpy(), t2n(layer.gamma_noise).reshape(n, c), t2n(layer.beta_noise).reshape(n, c),
                      t2n(layer.lmda).reshape(n), {})
    x = cl(n2t(x_np)).requires_grad_(True)
    y = layer(x)
    if c > 1:
        assert y.is_contiguous(memory_format=torch.channels_last)      # the format is kept, like the reference does
    y.backward(cl(n2t(dy_np)))
    y64, cache = O.forward(x_np, st, dtype=np.float64)
    dx64, dg64, db64, dl64 = O.backward(dy_np, x_np, st, cache, dtype=np.float64)
    assert_rel(t2n(y), y64, FWD_RTOL, "y")
    assert_rel(t2n(x.grad), dx64, GRAD_RTOL, "dx")
    assert_rel(t2n(layer.gamma_noise.grad).reshape(n, c), dg64, GRAD_RTOL, "d_gamma")
    assert_rel(t2n(layer.beta_noise.grad).reshape(n, c), db64, GRAD_RTOL, "d_beta")
    assert_rel(t2n(layer.lmda.grad).reshape(n), dl64, GRAD_RTOL, "d_lmda", scale=max(np.abs(dl64).max(), 1e-3))


@pytest.mark.parametrize("shape", [(4, 64, 48, 48), (3, 24, 40, 40), (4, 5, 17, 19)])
def test_nhwc_instance_stats_hard_planes(shape):
    """Kernel 1 (NHWC) through the C ABI with a large-mean / tiny-sigma channel, twice (workspace left zeroed,
    run-to-run deterministic)."""
    from maxstyle_b200 import functional as F, _lib as L
    n, c, h, w = shape
    x_np = make_input(11, shape)
    x_np[0, 1] = x_np[0, 1] * 1e-3 + 50.0
    x = cl(n2t(x_np))
    ws = F.new_workspace(n, c, h, w, L.F32, x.device, L.NHWC)
    mu, sig = F.instance_stats(x, 1e-6, ws)
    mu64, sig64 = O.instance_stats(x_np, 1e-6, dtype=np.float64)
    assert_rel(t2n(mu), mu64, 1e-6, "mu")
    assert np.abs(t2n(sig) / sig64 - 1).max() < 1e-5
    mu2, sig2 = F.instance_stats(x, 1e-6, ws)
    assert torch.equal(mu, mu2) and torch.equal(sig, sig2)


@pytest.mark.parametrize("shape", [(4, 32, 40, 40), (3, 16, 33, 31)])
def test_nhwc_bf16_matches_reference_on_float_input(shape):
    """bf16 storage, fp32 arithmetic (documented extension): oracle = the reference on x.float()."""
    n, c, h, w = shape
    torch.manual_seed(5)
    layer = make_layer(n, c)
    x_np = make_input(5, shape)
    xb = cl(n2t(x_np).to(torch.bfloat16))
    dyb = cl(n2t(np.random.RandomState(6).standard_normal(size=shape).astype(np.float32)).to(torch.bfloat16))
    st = oracle_state(layer.perm.numpy(), t2n(layer.gamma_noise).reshape(n, c), t2n(layer.beta_noise).reshape(n, c),
                      t2n(layer.lmda).reshape(n), {})
    x = xb.clone().requires_grad_(True)
    y = layer(x)
    assert y.dtype == torch.bfloat16
    y.backward(dyb)
    xf, dyf = t2n(xb), t2n(dyb)
    y64, cache = O.forward(xf, st, dtype=np.float64)
    dx64, dg64, db64, dl64 = O.backward(dyf, xf, st, cache, dtype=np.float64)
    assert_rel(t2n(y), y64, 2 ** -8, "y (bf16 rounding of the output)")
    assert_rel(t2n(x.grad), dx64, 2 ** -8, "dx (bf16 rounding of the output)")
    assert_rel(t2n(layer.gamma_noise.grad).reshape(n, c), dg64, GRAD_RTOL, "d_gamma")
    assert_rel(t2n(layer.beta_noise.grad).reshape(n, c), db64, GRAD_RTOL, "d_beta")
    assert_rel(t2n(layer.lmda.grad).reshape(n), dl64, GRAD_RTOL, "d_lmda", scale=max(np.abs(dl64).max(), 1e-3))


def test_nhwc_agrees_with_nchw_kernels_and_fused_adam():
    """Same logical tensor through both layouts: same tables (to rounding), same Adam trajectory."""
    from maxstyle_b200 import FusedStyleOptimizer
    n, c, h, w = 6, 32, 28, 28
    x_np = make_input(3, (n, c, h, w))
    dy_np = np.random.RandomState(4).standard_normal(size=(n, c, h, w)).astype(np.float32)
    outs = []
    for to_layout in (lambda t: t, cl):
        torch.manual_seed(9)
        layer = make_layer(n, c)
        opt = FusedStyleOptimizer([layer], lr=0.1)
        x = to_layout(n2t(x_np))
        dy = to_layout(n2t(dy_np))
        for _ in range(3):
            y = layer(x)
            y.backward(dy)
            opt.step()
        outs.append((t2n(y), t2n(layer.gamma_noise), t2n(layer.beta_noise), t2n(layer.lmda), opt.step_count(layer)))
    a, b = outs
    assert a[4] == b[4] == 3
    assert_rel(b[0], a[0], 1e-5, "y after 3 steps")
    for i, name in ((1, "gamma_noise"), (2, "beta_noise"), (3, "lmda")):
        assert_rel(b[i], a[i], 1e-4, name)


def test_nhwc_rejects_too_many_channel_vectors():
    from maxstyle_b200 import functional as F, _lib as L
    assert F.workspace_bytes(2, 4096, 4, 4, L.F32, L.NHWC) == 0        # 512 vectors per pixel: no kernel
    assert F.workspace_bytes(2, 2048, 4, 4, L.F32, L.NHWC) > 0
