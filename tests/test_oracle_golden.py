"""Pin the numpy oracle against the reference-generated golden vectors (tests/golden/,
made by oracle/gen_golden.py from the unmodified reference).  CPU only."""
import numpy as np
import pytest

from oracle import maxstyle_oracle as O
from oracle.gen_golden import make_input, FWD_BWD_CASES


def _state(g, pre, kwargs, p=1.0):
    return O.StyleState(perm=g[pre + "perm"], gamma_noise=g[pre + "gamma_noise"], beta_noise=g[pre + "beta_noise"],
                        lmda=g[pre + "lmda"], rand_p=float(g[pre + "rand_p"]), p=p,
                        mix_style=kwargs.get("mix_style", True), no_noise=kwargs.get("no_noise", False))


def _close(a, b, rtol, name, scale=None):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    s = np.abs(b).max() if scale is None else scale
    err = np.abs(a - b).max() if a.size else 0.0
    assert err <= rtol * max(s, 1e-30), f"{name}: max abs err {err:.3e} vs scale {s:.3e} (rtol {rtol})"


@pytest.mark.parametrize("idx", range(len(FWD_BWD_CASES)))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_forward_backward_matches_reference(golden, manifest, idx, dtype):
    g = golden["fwd_bwd"]
    meta = manifest["fwd_bwd"][idx]
    pre = f"f{idx}_"
    shape = (meta["N"], meta["C"], meta["H"], meta["W"])
    x = make_input(meta["seed"], shape, meta["kind"])
    dy = np.random.RandomState(meta["seed"] + 5000).standard_normal(size=shape).astype(np.float32)
    st = _state(g, pre, meta["kwargs"])
    y, cache = O.forward(x, st, dtype=dtype)
    # mean of an 'offset' plane carries fp32 summation error ~1e-7*|mu|; sig is then compared
    # against the scale of the data, everything else against its own max.
    _close(cache.mu, g[pre + "mu"], 1e-6, "mu")
    _close(cache.sig, g[pre + "sig"], 2e-5 if meta["kind"] == "offset" else 1e-5, "sig")
    _close(cache.gamma_std, g[pre + "gamma_std"], 1e-4 if meta["kind"] == "offset" else 1e-5, "gamma_std",
           scale=np.abs(g[pre + "sig"]).max())
    _close(cache.beta_std, g[pre + "beta_std"], 1e-5, "beta_std", scale=np.abs(g[pre + "mu"]).max())
    ytol = 2e-4 if meta["kind"] == "offset" else 1e-5       # offset: (x-mu)/sig amplifies mean rounding by |mu|/sig=1e4
    _close(y, g[pre + "y"], ytol, "y")
    dx, dgam, dbet, dlm = O.backward(dy, x, st, cache, dtype=dtype)
    # offset: |mu|/sig = 1e4, so the fp32 rounding of the reference's own mean (6e-8*|mu|) moves
    # x_normed by ~1e-3 relative; gradients that sum g*x_normed inherit that conditioning.
    gtol = 5e-3 if meta["kind"] == "offset" else 1e-4
    _close(dx, g[pre + "dx"], gtol, "dx")
    if g[pre + "d_gamma_noise"].size:
        _close(dgam, g[pre + "d_gamma_noise"], gtol, "d_gamma")
        _close(dbet, g[pre + "d_beta_noise"], gtol, "d_beta")
    if g[pre + "d_lmda"].size:
        # d_lmda multiplies by (mu[perm]-mu), itself a difference of two O(100) numbers in 'offset'
        _close(dlm, g[pre + "d_lmda"].reshape(-1), 2e-2 if meta["kind"] == "offset" else gtol, "d_lmda",
               scale=max(np.abs(g[pre + "d_lmda"]).max(), np.abs(dlm).max(), 1e-3))


def test_lmda_edges_mask_is_inclusive(golden, manifest):
    idx = [m["name"] for m in manifest["fwd_bwd"]].index("lmda_edges")
    g = golden["fwd_bwd"]
    d = g[f"f{idx}_d_lmda"].reshape(-1)
    lm = g[f"f{idx}_lmda"]
    assert d[1] == 0.0 and d[3] == 0.0          # outside [0,1]: clamp kills the gradient
    assert d[0] != 0.0 and d[2] != 0.0          # exactly 0 and exactly 1 still receive gradient
    assert lm[0] == 0.0 and lm[2] == 1.0


def test_gamma_std_is_cached(golden, manifest):
    g = golden["cache"]; meta = manifest["cache"]
    shape = (meta["N"], meta["C"], meta["H"], meta["W"])
    st = _state(g, "k_", {})
    x1 = make_input(meta["seed"], shape)
    x2 = make_input(meta["seed2"], shape) * np.float32(meta["scale2"])
    y1, _ = O.forward(x1, st)
    gs = st.gamma_std.copy()
    y2, _ = O.forward(x2, st)
    assert np.array_equal(gs, st.gamma_std)
    _close(y1, g["k_y1"], 1e-5, "y1"); _close(y2, g["k_y2"], 1e-5, "y2")
    _close(st.gamma_std, g["k_gamma_std"], 1e-5, "gamma_std")


def test_reference_selftest_trajectory(golden):
    """maxstyle.py:193-241 (seed 43): forward + closed-form backward + Adam restatement
    reproduce the five printed losses and the parameter trajectory."""
    g = golden["selftest"]
    x = (3 * np.arange(32, dtype=np.float32) + 5).reshape(4, 2, 2, 2)
    st = _state(g, "s_", {}, p=0.5)
    assert list(st.perm) == [0, 1, 3, 2] and abs(st.rand_p - 0.34617907) < 1e-7
    adam = {k: O.AdamState(np.zeros_like(getattr(st, k)), np.zeros_like(getattr(st, k)))
            for k in ("gamma_noise", "beta_noise", "lmda")}
    for i in range(5):
        y, cache = O.forward(x, st)
        loss = np.mean((y - 1.0) ** 2, dtype=np.float64)
        assert abs(loss - g["s_losses"][i]) <= 1e-5 * g["s_losses"][i]
        dy = (2.0 * (y - 1.0) / y.size).astype(np.float32)
        _, dgam, dbet, dlm = O.backward(dy, x, st, cache)
        st.gamma_noise = O.adam_step(st.gamma_noise, dgam, adam["gamma_noise"])
        st.beta_noise = O.adam_step(st.beta_noise, dbet, adam["beta_noise"])
        st.lmda = O.adam_step(st.lmda, dlm, adam["lmda"])
        _close(st.gamma_noise, g[f"s_step{i}_gamma_noise"], 1e-5, "gamma")
        _close(st.beta_noise, g[f"s_step{i}_beta_noise"], 1e-5, "beta")
        _close(st.lmda, g[f"s_step{i}_lmda"].reshape(-1), 1e-5, "lmda")
    assert np.all(st.gamma_std == 0) and abs(st.beta_std[0] - 30.98386765) < 1e-4


def test_adam_and_grads_on_loop_fixture(golden, manifest):
    """Given the recorded gradients, the Adam restatement reproduces torch.optim.Adam's
    parameter trajectory (advanced_triplet_recon_segmentation_model.py:537,562)."""
    g = golden["loop"]
    params = {k: g["l_" + k].copy() for k in ("gamma_noise", "beta_noise", "lmda")}
    adam = {k: O.AdamState(np.zeros_like(v), np.zeros_like(v)) for k, v in params.items()}
    for i in range(5):
        for k in params:
            grad = g[f"l_step{i}_grad_{k}"].reshape(params[k].shape)
            params[k] = O.adam_step(params[k], grad, adam[k])
            _close(params[k], g[f"l_step{i}_{k}"].reshape(params[k].shape), 2e-6, f"step{i} {k}")


def test_identity_cases():
    st = O.StyleState(perm=np.array([1, 0]), gamma_noise=np.zeros((2, 1)), beta_noise=np.zeros((2, 1)),
                      lmda=np.zeros(2), rand_p=0.7, p=0.5)
    x = np.ones((2, 1, 4, 4), np.float32)
    y, c = O.forward(x, st)
    assert y is x and c.identity
    st.rand_p = 0.1
    assert O.forward(np.ones((1, 1, 4, 4), np.float32), st)[0] is not None
    assert O.is_identity_case(st, (1, 1, 4, 4)) and O.is_identity_case(st, (2, 1, 1, 1))
    st.mix_style, st.no_noise = False, True
    assert O.is_identity_case(st, (2, 1, 4, 4))


def test_global_batch_extension_equals_concatenated_reference(golden, manifest):
    """Sharded evaluation (rows of a global batch + gathered mu/sig tables) equals the
    reference on the concatenated batch."""
    idx = [m["name"] for m in manifest["fwd_bwd"]].index("mid_16ch")
    meta = manifest["fwd_bwd"][idx]; g = golden["fwd_bwd"]; pre = f"f{idx}_"
    shape = (meta["N"], meta["C"], meta["H"], meta["W"])
    x = make_input(meta["seed"], shape, meta["kind"])
    dy = np.random.RandomState(meta["seed"] + 5000).standard_normal(size=shape).astype(np.float32)
    gmu, gsig = O.instance_stats(x, 1e-6)
    R = 2; nl = shape[0] // R
    for r in range(R):
        rows = slice(r * nl, (r + 1) * nl)
        st = O.StyleState(perm=g[pre + "perm"], gamma_noise=g[pre + "gamma_noise"][rows],
                          beta_noise=g[pre + "beta_noise"][rows], lmda=g[pre + "lmda"][rows], p=1.0)
        y, cache = O.forward(x[rows], st, global_mu=gmu, global_sig=gsig, row_offset=r * nl)
        _close(y, g[pre + "y"][rows], 1e-5, "y shard")
        dx, dgam, dbet, dlm = O.backward(dy[rows], x[rows], st, cache, global_mu=gmu, global_sig=gsig,
                                         row_offset=r * nl)
        _close(dx, g[pre + "dx"][rows], 1e-4, "dx shard")
        _close(dgam, g[pre + "d_gamma_noise"][rows], 1e-4, "dgamma shard")
        _close(dlm, g[pre + "d_lmda"].reshape(-1)[rows], 1e-4, "dlmda shard",
               scale=np.abs(g[pre + "d_lmda"]).max())


def test_torch_port_matches_goldens(golden, manifest):
    """The CPU timing baseline (oracle/torch_port.py) computes what the reference computes."""
    import torch
    from oracle.torch_port import StylePort
    torch.set_num_threads(1)
    g = golden["fwd_bwd"]
    for idx, meta in enumerate(manifest["fwd_bwd"]):
        if meta["kind"] == "offset":
            continue
        kw = meta["kwargs"]; pre = f"f{idx}_"
        shape = (meta["N"], meta["C"], meta["H"], meta["W"])
        port = StylePort(g[pre + "perm"], g[pre + "gamma_noise"], g[pre + "beta_noise"], g[pre + "lmda"],
                         mix_style=kw.get("mix_style", True), no_noise=kw.get("no_noise", False))
        x = torch.from_numpy(make_input(meta["seed"], shape, meta["kind"])).requires_grad_(True)
        dy = torch.from_numpy(np.random.RandomState(meta["seed"] + 5000).standard_normal(size=shape).astype(np.float32))
        y = port.forward(x)
        y.backward(dy)
        _close(y.detach().numpy(), g[pre + "y"], 1e-6, "port y")
        _close(x.grad.numpy(), g[pre + "dx"], 1e-6, "port dx")
        if g[pre + "d_lmda"].size and kw.get("mix_learnable", True):
            _close(port.lmda.grad.numpy().reshape(-1), g[pre + "d_lmda"].reshape(-1), 1e-5, "port d_lmda",
                   scale=max(np.abs(g[pre + "d_lmda"]).max(), 1e-3))


@pytest.mark.parametrize("idx", range(len(FWD_BWD_CASES)))
def test_c_oracle_matches_reference_and_numpy_oracle(golden, manifest, idx):
    """The C restatement (oracle/maxstyle_oracle.c, double precision, scalar loops) against the reference-generated goldens
    and against the numpy oracle: two independently written checkers must agree to float64 rounding."""
    from oracle import build_c as CO
    g = golden["fwd_bwd"]
    meta = manifest["fwd_bwd"][idx]
    pre = f"f{idx}_"
    shape = (meta["N"], meta["C"], meta["H"], meta["W"])
    x = make_input(meta["seed"], shape, meta["kind"])
    dy = np.random.RandomState(meta["seed"] + 5000).standard_normal(size=shape).astype(np.float32)
    st = _state(g, pre, meta["kwargs"])
    if O.is_identity_case(st, shape):
        pytest.skip("identity case: the layer returns x itself")
    flags = CO.flags_of(mix_style=st.mix_style, no_noise=st.no_noise, compute_std=True)
    y, cache = CO.forward(x, st.perm, st.lmda, st.gamma_noise, st.beta_noise, eps=st.eps, flags=flags)
    dx, dgam, dbet, dlm = CO.backward(dy, x, cache)
    # against the numpy oracle in float64: same algorithm, different code
    st2 = _state(g, pre, meta["kwargs"])
    y64, c64 = O.forward(x, st2, dtype=np.float64)
    dx64, dg64, db64, dl64 = O.backward(dy, x, st2, c64, dtype=np.float64)
    for name, a, b in (("y", y, y64), ("mu", cache["mu"], c64.mu), ("sig", cache["sig"], c64.sig), ("dx", dx, dx64),
                       ("d_gamma", dgam, dg64), ("d_beta", dbet, db64), ("d_lmda", dlm, dl64),
                       ("gamma_std", cache["gamma_std"], c64.gamma_std), ("beta_std", cache["beta_std"], c64.beta_std)):
        _close(a, b, 1e-9, f"C vs numpy oracle: {name}", scale=max(np.abs(b).max(), 1e-6))
    # against the reference's own fp32 outputs
    off = meta["kind"] == "offset"
    _close(y, g[pre + "y"], 2e-4 if off else 1e-5, "y")
    _close(dx, g[pre + "dx"], 5e-3 if off else 1e-4, "dx")
    _close(cache["sig"], g[pre + "sig"], 2e-5 if off else 1e-5, "sig")
