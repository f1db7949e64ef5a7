"""Constructor / RNG contract of the drop-in module against the reference-generated goldens
(reference: src/advanced/maxstyle.py:14-122).  CPU only: constructs with use_gpu=False."""
import numpy as np
import pytest
import torch

import maxstyle_b200
from maxstyle_b200 import MaxStyle

KW = ("mix_style", "no_noise", "mix_learnable", "noise_learnable", "always_use_beta")


def _np(t):
    return t.detach().cpu().numpy()


def test_all_ctor_cases_match_reference(golden, manifest):
    g = golden["ctor"]
    checked = raised = 0
    for e in manifest["ctor"]:
        kw = {k: e[k] for k in KW}
        torch.manual_seed(e["seed"])
        if e["raises"]:
            with pytest.raises(AssertionError):
                MaxStyle(e["N"], e["C"], p=e["p"], use_gpu=False, **kw)
            raised += 1
            continue
        m = MaxStyle(e["N"], e["C"], p=e["p"], use_gpu=False, **kw)
        pre = f"c{e['idx']}_"
        assert m.perm.dtype == torch.int64 and np.array_equal(_np(m.perm), g[pre + "perm"]), e      # bit-exact
        assert np.float32(m.rand_p.item()) == g[pre + "rand_p"], e
        for name in ("gamma_noise", "beta_noise", "lmda"):
            t = getattr(m, name)
            want = g[pre + name]
            assert tuple(t.shape) == ((e["N"], e["C"], 1, 1) if name != "lmda" else (e["N"], 1, 1, 1))
            assert np.array_equal(_np(t).reshape(want.shape), want), (e, name)                       # same draws
            short = name.split("_")[0]
            assert isinstance(t, torch.nn.Parameter) == bool(g[pre + short + "_is_param"]), (e, name)
            assert t.requires_grad == bool(g[pre + short + "_requires_grad"]), (e, name)
        assert [k for k, _ in m.named_parameters()] == e["param_names"]
        assert list(m.state_dict().keys()) == e["state_dict_keys"]
        assert len(list(m.parameters())) == e["n_params"]
        # the constructor consumed exactly as much of the CPU generator as the reference
        assert float(torch.rand(1).item()) == e["next_cpu_rand"], e
        assert m.gamma_std is None and m.beta_std is None and m.data is None
        checked += 1
    assert checked + raised == len(manifest["ctor"]) and checked >= 150 and raised >= 10


def test_signature_and_attributes():
    import inspect
    sig = inspect.signature(MaxStyle.__init__)
    assert list(sig.parameters)[1:] == ["batch_size", "num_feature", "p", "mix_style", "no_noise", "mix_learnable",
                                        "noise_learnable", "always_use_beta", "alpha", "eps", "use_gpu", "debug"]
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect._empty}
    assert d == dict(p=0.5, mix_style=True, no_noise=False, mix_learnable=True, noise_learnable=True,
                     always_use_beta=False, alpha=0.1, eps=1e-6, use_gpu=True, debug=False)
    torch.manual_seed(3)
    m = MaxStyle(5, 2, p=1.0, use_gpu=False)
    for attr in ("batch_size", "num_feature", "p", "mix_style", "no_noise", "mix_learnable", "noise_learnable",
                 "always_use_beta", "alpha", "eps", "use_gpu", "debug", "device", "data", "perm", "rand_p",
                 "gamma_std", "beta_std", "gamma_noise", "beta_noise", "lmda"):
        assert hasattr(m, attr), attr
    assert m.device == torch.device("cpu")
    assert repr(m).startswith("MaxStyle:") and "mean of mix coefficient" in repr(m)
    # usable in the reference's container and optimiser (model:527,537)
    d = torch.nn.ModuleDict({"3": m})
    opt = torch.optim.Adam(d.parameters(), lr=0.1)
    assert len(opt.param_groups[0]["params"]) == 3
    d.zero_grad()


def test_inactive_module_is_identity_and_has_no_parameters():
    torch.manual_seed(0)
    m = MaxStyle(4, 3, p=0.0, use_gpu=False)           # rand_p >= 0 always: never active
    assert list(m.parameters()) == [] and repr(m) == "diffuse style not applied"
    x = torch.randn(4, 3, 5, 5)
    assert m(x) is x and m.data is x                    # same object, even on CPU (maxstyle.py:146-152)


def test_early_outs_return_same_object_without_touching_cuda():
    torch.manual_seed(1)
    m = MaxStyle(1, 3, p=1.0, use_gpu=False)
    x1 = torch.randn(1, 3, 4, 4)
    assert m(x1) is x1                                  # B <= 1
    m = MaxStyle(4, 3, p=1.0, use_gpu=False)
    x2 = torch.randn(4, 3, 1, 1)
    assert m(x2) is x2                                  # spatial size 1
    m = MaxStyle(4, 3, p=1.0, mix_style=False, no_noise=True, noise_learnable=False, use_gpu=False)
    x3 = torch.randn(4, 3, 4, 4)
    assert m(x3) is x3                                  # nothing to do


def test_shape_mismatch_asserts_and_cpu_forward_fails_loudly():
    torch.manual_seed(2)
    m = MaxStyle(4, 3, p=1.0, use_gpu=False)
    with pytest.raises(AssertionError):
        m(torch.randn(4, 2, 5, 5))
    with pytest.raises(AssertionError):
        m(torch.randn(3, 3, 5, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(4, 3, 5, 5))


def test_reset_redraws_and_drops_cache():
    torch.manual_seed(5)
    m = MaxStyle(6, 2, p=1.0, use_gpu=False)
    old = m.lmda
    m.gamma_std = torch.ones(1, 2, 1, 1)
    m.reset()
    assert m.gamma_std is None and m.beta_std is None
    assert m.lmda is not old and [k for k, _ in m.named_parameters()] == ["gamma_noise", "beta_noise", "lmda"]
    # active -> inactive -> active transitions keep the parameter registry consistent
    m.p = 0.0
    m.reset()
    assert list(m.parameters()) == [] and not isinstance(m.lmda, torch.nn.Parameter)
    m.p = 1.0
    m.reset()
    assert [k for k, _ in m.named_parameters()] == ["gamma_noise", "beta_noise", "lmda"]


def test_product_package_never_imports_the_oracle():
    import os, re
    pkg = os.path.dirname(maxstyle_b200.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "/root/reference" not in src, f


@pytest.mark.parametrize("kw", [dict(), dict(always_use_beta=True), dict(no_noise=True, noise_learnable=False),
                                dict(mix_style=False), dict(mix_learnable=False)])
def test_reinit_in_place_draws_like_a_fresh_construction(kw):
    """MaxStyle.reinit_() (used by StyleLoopExecutor for CUDA-graph replay) must consume the generators exactly like
    init_parameters() and write the same values -- into the SAME tensors."""
    import torch
    from maxstyle_b200 import MaxStyle
    n, c = 7, 5
    torch.manual_seed(99)
    m = MaxStyle(n, c, p=2.0, use_gpu=False, **kw)       # p = 2: storage for the active case
    m.p = 0.5
    ptrs = [t.data_ptr() for t in (m.gamma_noise, m.beta_noise, m.lmda)]
    for seed in range(8):
        torch.manual_seed(seed)
        active = m.reinit_()
        after = float(torch.rand(1))
        torch.manual_seed(seed)
        f = MaxStyle(n, c, p=0.5, use_gpu=False, **kw)
        assert float(torch.rand(1)) == after, "generator consumed differently"
        assert torch.equal(m.perm, f.perm) and torch.equal(m.rand_p, f.rand_p)
        assert active == bool(f.rand_p < f.p)
        if active:
            if kw.get("noise_learnable", True):
                assert torch.equal(m.gamma_noise.detach(), f.gamma_noise.detach()) and torch.equal(m.beta_noise.detach(), f.beta_noise.detach())
            if kw.get("mix_style", True):
                assert torch.equal(m.lmda.detach(), f.lmda.detach())
        assert [t.data_ptr() for t in (m.gamma_noise, m.beta_noise, m.lmda)] == ptrs, "reinit_ must not re-allocate"
