"""The single-kernel forwards built on {value, tag} words -- the paired forward (csrc/pair_fwd.cuh: a CTA owns a piece of a
plane for both passes) and the cluster-resident forward (csrc/cluster_fwd.cuh: planes held in shared memory across a
thread-block cluster) -- against the two-pass path and the float64 oracle: every piece count / cluster size / stage count
they can run with, planes that do not split evenly, the cycle-ordered walk for batches larger than the grid, flag variants,
permutations with fixed points, and replays (the launch tag must advance)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import maxstyle_oracle as O
from oracle.gen_golden import make_input
from test_gpu_parity import FWD_RTOL, assert_rel, make_layer, n2t, oracle_state, t2n

pytestmark = pytest.mark.gpu

NAMES = ("y", "mu", "sig", "scale", "shift", "gamma_std", "beta_std")


def sweep_for(cs=0, stages=0, pieces=0, force=True):
    """cs == -1: the paired forward with `pieces` pieces per plane; otherwise the cluster-resident forward."""
    from maxstyle_b200 import _lib as L
    if cs < 0:
        return (L.SWEEP_FORCE_PAIR if force else 0) | (pieces << L.SWEEP_CLUSTER_PIECES_SHIFT)
    return ((L.SWEEP_FORCE_CLUSTER if force else 0) | (cs << L.SWEEP_CLUSTER_SIZE_SHIFT) | (stages << L.SWEEP_CLUSTER_STAGES_SHIFT)
            | (pieces << L.SWEEP_CLUSTER_PIECES_SHIFT))


def run_fwd(x, perm, lmda, gamma, beta, gs, bs, flags, ws, sweep):
    from maxstyle_b200 import functional as F
    n, c, h, w = x.shape
    old = F.SWEEP_STATS
    F.SWEEP_STATS = sweep
    try:
        out = F.forward_raw(x, perm, lmda, gamma, beta, gs, bs, flags, 1e-6, ws)
    finally:
        F.SWEEP_STATS = old
    F.workspace_status(ws, n, c, h, w, F.dtype_code(x))
    return [t.clone() for t in out] + [gs, bs]


SHAPES = [
    (20, 64, 224, 224, torch.float32),    # config 1: 196 KB planes
    (20, 16, 96, 96, torch.float32),      # 36 KB planes: one CTA holds five of them
    (32, 16, 192, 192, torch.float32),    # config 3
    (20, 1, 224, 224, torch.float32),     # C = 1
    (3, 2, 160, 160, torch.float32),      # fewer items than SMs
    (2, 3, 130, 130, torch.float32),      # 67600-byte planes: parts of unequal size
    (6, 8, 128, 128, torch.bfloat16),
    (2, 2, 512, 512, torch.float32),      # config-5 planes (1 MiB): only a cluster holds one
    (5, 3, 100, 84, torch.bfloat16),      # 16800-byte planes, ragged chunks
]


GEOMETRIES = [(-1, 0, 0), (-1, 0, 1), (-1, 0, 2), (-1, 0, 3), (-1, 0, 4), (-1, 0, 7), (-1, 0, 16), (-1, 0, 32),   # paired: pieces
              (0, 0, 0), (1, 0, 1), (2, 0, 1), (4, 0, 1), (8, 0, 1), (2, 1, 1),       # (CTAs per cluster, stage cap, pieces)
              (1, 0, 4), (2, 0, 2), (2, 1, 2), (1, 2, 3), (2, 0, 8)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("cs,stages,pieces", GEOMETRIES)
def test_cluster_forward_matches_two_pass_and_oracle(shape, cs, stages, pieces):
    from maxstyle_b200 import functional as F, _lib as L
    n, c, h, w, dt = shape
    lib = L.get_lib()
    torch.manual_seed(n * 7 + c)
    x = n2t(make_input(11 + n + c + h, (n, c, h, w)), dt)
    code = F.dtype_code(x)
    sweep = sweep_for(cs, stages, pieces)
    if lib.maxstyle_fwd_kernels(n, c, h, w, code, L.NCHW, sweep) != 1:
        pytest.skip("shape does not run with this cluster geometry")
    layer = make_layer(n, c)
    ws = F.new_workspace(n, c, h, w, code, x.device)
    perm = layer.perm.to(x.device)
    lm, gn, bn = layer.lmda.detach(), layer.gamma_noise.detach(), layer.beta_noise.detach()
    z = lambda: torch.zeros(c, device=x.device)
    first = L.FLAG_MIX_STYLE | L.FLAG_COMPUTE_BATCH_STD
    two = run_fwd(x, perm, lm, gn, bn, z(), z(), first, ws, L.SWEEP_NO_FUSED)
    a = run_fwd(x, perm, lm, gn, bn, z(), z(), first, ws, sweep)
    b = run_fwd(x, perm, lm, gn, bn, z(), z(), first, ws, sweep)
    gs0, bs0 = two[5], two[6]
    two_c = run_fwd(x, perm, lm, gn, bn, gs0.clone(), bs0.clone(), L.FLAG_MIX_STYLE, ws, L.SWEEP_NO_FUSED)
    cached = run_fwd(x, perm, lm, gn, bn, gs0.clone(), bs0.clone(), L.FLAG_MIX_STYLE, ws, sweep)
    tol = 2.0 ** -8 if dt == torch.bfloat16 else 2e-6
    for cand, ref, what in ((a, two, "first"), (cached, two_c, "cached")):
        for i, nm in enumerate(NAMES):
            r = t2n(ref[i])
            assert_rel(t2n(cand[i]), r, tol if nm == "y" else 1e-5, f"{what} {nm}", scale=max(np.abs(r).max(), 1e-3))
    for i, nm in enumerate(NAMES):
        assert torch.equal(a[i], b[i]), f"{nm}: not deterministic"
    st = oracle_state(layer.perm.numpy(), t2n(gn).reshape(n, c), t2n(bn).reshape(n, c), t2n(lm).reshape(n), {})
    y64, cache = O.forward(t2n(x), st, dtype=np.float64)
    for cand in (a, cached):
        assert_rel(t2n(cand[0]), y64, tol if dt == torch.bfloat16 else FWD_RTOL, "y vs f64 oracle")
        assert_rel(t2n(cand[1]), cache.mu, 1e-6, "mu", scale=max(np.abs(cache.mu).max(), 1e-3))
        assert np.abs(t2n(cand[2]) / cache.sig - 1).max() < 1e-5


@pytest.mark.parametrize("shape,cs,pieces", [((200, 2, 64, 64), 1, 1), ((300, 1, 72, 64), 1, 1), ((100, 2, 128, 128), 1, 4),
                                             ((60, 3, 160, 128), 2, 2), ((700, 1, 64, 64), -1, 1), ((100, 2, 128, 128), -1, 8),
                                             ((40, 2, 256, 256), -1, 16)])
def test_cluster_forward_cycle_order_when_batch_exceeds_clusters(shape, cs, pieces):
    """N above the number of co-resident clusters: the cached-std forward walks the samples in cycle order of perm (an item
    waits only for the next one); the first forward of a module (whole-channel dependency) takes another path."""
    from maxstyle_b200 import functional as F, _lib as L
    n, c, h, w = shape
    torch.manual_seed(n)
    x = n2t(make_input(3 + n, (n, c, h, w)))
    layer = make_layer(n, c)
    ws = F.new_workspace(n, c, h, w, L.F32, x.device)
    perm = layer.perm.to(x.device)
    lm, gn, bn = layer.lmda.detach(), layer.gamma_noise.detach(), layer.beta_noise.detach()
    z = lambda: torch.zeros(c, device=x.device)
    first = L.FLAG_MIX_STYLE | L.FLAG_COMPUTE_BATCH_STD
    sweep = sweep_for(cs, 0, pieces)
    geo = (C.c_int * 12)()
    assert L.get_lib().maxstyle_fwd_geometry(n, c, h, w, L.F32, sweep, geo) == 0 and geo[9] == 1, "expected the cycle-ordered walk"
    two = run_fwd(x, perm, lm, gn, bn, z(), z(), first, ws, L.SWEEP_NO_FUSED)
    dflt = run_fwd(x, perm, lm, gn, bn, z(), z(), first, ws, sweep)               # declines, falls through to another path
    two_c = run_fwd(x, perm, lm, gn, bn, two[5].clone(), two[6].clone(), L.FLAG_MIX_STYLE, ws, L.SWEEP_NO_FUSED)
    for rep in range(3):
        cached = run_fwd(x, perm, lm, gn, bn, two[5].clone(), two[6].clone(), L.FLAG_MIX_STYLE, ws, sweep)
        for i, nm in enumerate(NAMES):
            r = t2n(two_c[i])
            assert_rel(t2n(cached[i]), r, 2e-6 if nm == "y" else 1e-5, f"cached {nm} rep {rep}", scale=max(np.abs(r).max(), 1e-3))
    for i, nm in enumerate(NAMES):
        r = t2n(two[i])
        assert_rel(t2n(dflt[i]), r, 2e-6 if nm == "y" else 1e-5, f"first {nm}", scale=max(np.abs(r).max(), 1e-3))


@pytest.mark.parametrize("cs,pieces", [(-1, 1), (-1, 2), (1, 1), (2, 1), (4, 1), (1, 2)])
def test_cluster_forward_flag_variants_and_fixed_points(cs, pieces):
    from maxstyle_b200 import functional as F, _lib as L
    n, c, h, w = 6, 4, 96, 96
    x = n2t(make_input(5, (n, c, h, w)))
    ws = F.new_workspace(n, c, h, w, L.F32, x.device)
    g = torch.Generator().manual_seed(3)
    gamma, beta = torch.randn(n, c, generator=g).cuda(), torch.randn(n, c, generator=g).cuda()
    lmda = (torch.rand(n, generator=g) * 1.6 - 0.3).cuda()                   # some outside [0, 1]
    perm = torch.tensor([0, 2, 1, 3, 5, 4], dtype=torch.int64).cuda()      # fixed points 0 and 3
    for flags in (L.FLAG_MIX_STYLE | L.FLAG_COMPUTE_BATCH_STD, L.FLAG_MIX_STYLE | L.FLAG_NO_NOISE, L.FLAG_COMPUTE_BATCH_STD,
                  L.FLAG_MIX_STYLE | L.FLAG_NO_NOISE | L.FLAG_NO_CLAMP, L.FLAG_MIX_STYLE | L.FLAG_NO_NOISE | L.FLAG_COMPUTE_BATCH_STD):
        z = lambda: torch.zeros(c, device=x.device)
        two = run_fwd(x, perm, lmda, gamma, beta, z(), z(), flags, ws, L.SWEEP_NO_FUSED)
        clu = run_fwd(x, perm, lmda, gamma, beta, z(), z(), flags, ws, sweep_for(cs, 0, pieces))
        for i, nm in enumerate(NAMES):
            r = t2n(two[i])
            assert_rel(t2n(clu[i]), r, 2e-6 if nm == "y" else 1e-5, f"flags {flags} {nm}", scale=max(np.abs(r).max(), 1e-3))


def test_cluster_forward_replays_with_new_data():
    """Twenty launches on one workspace with x changing in between: every launch must see ITS statistics (the {value, tag}
    words of earlier launches stay in the table with older tags)."""
    from maxstyle_b200 import functional as F, _lib as L
    n, c, h, w = 20, 8, 96, 96
    layer = make_layer(n, c)
    dev = torch.device("cuda:0")
    ws = F.new_workspace(n, c, h, w, L.F32, dev)
    perm = layer.perm.to(dev)
    lm, gn, bn = layer.lmda.detach(), layer.gamma_noise.detach(), layer.beta_noise.detach()
    gs, bs = torch.rand(c, device=dev) + 0.5, torch.rand(c, device=dev) + 0.5
    for rep in range(20):
        x = n2t(make_input(100 + rep, (n, c, h, w))) * (1.0 + rep)
        two = run_fwd(x, perm, lm, gn, bn, gs.clone(), bs.clone(), L.FLAG_MIX_STYLE, ws, L.SWEEP_NO_FUSED)
        clu = run_fwd(x, perm, lm, gn, bn, gs.clone(), bs.clone(), L.FLAG_MIX_STYLE, ws, sweep_for(-1 if rep % 2 else 0))
        for i, nm in enumerate(NAMES[:5]):
            r = t2n(two[i])
            assert_rel(t2n(clu[i]), r, 2e-6 if nm == "y" else 1e-5, f"rep {rep} {nm}", scale=max(np.abs(r).max(), 1e-3))
