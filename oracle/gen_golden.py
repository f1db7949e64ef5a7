#!/usr/bin/env python
"""Generate golden input/output vectors from the UNMODIFIED reference layer.

Run in the build container only (the reference is mounted at /root/reference there and
nowhere else):

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz + MANIFEST.json

The reference module is imported by path (nothing is copied); every case runs
``MaxStyle(..., use_gpu=False)`` on CPU under a fixed ``torch.manual_seed`` and records
the tensors the parity tests need.  Inputs are regenerated from numpy seeds by
``make_input`` below so the fixtures only carry outputs and the random module state.
"""
from __future__ import annotations

import importlib.util
import itertools
import json
import os
import sys

import numpy as np
import torch

REF_FILE = "/root/reference/src/advanced/maxstyle.py"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference():
    spec = importlib.util.spec_from_file_location("_reference_maxstyle", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MaxStyle


def make_input(seed: int, shape, kind: str = "scaled") -> np.ndarray:
    """Deterministic synthetic feature map (numpy legacy RandomState is stable across
    versions/platforms).  'scaled' gives every plane its own scale/shift so the mu/sig
    tables and their batch std are non-degenerate."""
    rs = np.random.RandomState(seed)
    n, c = shape[0], shape[1]
    x = rs.standard_normal(size=shape).astype(np.float32)
    if kind == "scaled":
        scale = rs.uniform(0.5, 2.0, size=(n, c) + (1,) * (len(shape) - 2)).astype(np.float32)
        shift = rs.uniform(-1.0, 1.0, size=(n, c) + (1,) * (len(shape) - 2)).astype(np.float32)
        x = x * scale + shift
    elif kind == "sigmoid":       # layer-5-like: values in (0,1), small sigma
        x = (1.0 / (1.0 + np.exp(-x))).astype(np.float32)
    elif kind == "offset":        # large mean relative to sigma (cancellation stress)
        x = (x * 0.01 + 100.0).astype(np.float32)
    return x


def t2n(t):
    return t.detach().cpu().numpy().copy()      # copy: CPU tensors share memory with numpy views


def module_state(m):
    return dict(
        perm=t2n(m.perm).astype(np.int64),
        rand_p=np.float32(m.rand_p.item()),
        gamma_noise=t2n(m.gamma_noise).reshape(m.batch_size, m.num_feature),
        beta_noise=t2n(m.beta_noise).reshape(m.batch_size, m.num_feature),
        lmda=t2n(m.lmda).reshape(m.batch_size),
        gamma_is_param=np.bool_(isinstance(m.gamma_noise, torch.nn.Parameter)),
        beta_is_param=np.bool_(isinstance(m.beta_noise, torch.nn.Parameter)),
        lmda_is_param=np.bool_(isinstance(m.lmda, torch.nn.Parameter)),
        gamma_requires_grad=np.bool_(m.gamma_noise.requires_grad),
        beta_requires_grad=np.bool_(m.beta_noise.requires_grad),
        lmda_requires_grad=np.bool_(m.lmda.requires_grad),
    )


def gen_ctor_cases(MaxStyle):
    """Constructor / RNG contract (maxstyle.py:48-122) for all flag combinations."""
    out = {}
    meta = []
    idx = 0
    flags = list(itertools.product([True, False], repeat=5))
    for seed in (0, 1, 7, 43, 123):
        for mix_style, no_noise, mix_learnable, noise_learnable, always_use_beta in flags:
            kw = dict(mix_style=mix_style, no_noise=no_noise, mix_learnable=mix_learnable,
                      noise_learnable=noise_learnable, always_use_beta=always_use_beta)
            torch.manual_seed(seed)
            entry = dict(idx=idx, seed=seed, N=6, C=3, p=0.5, **kw)
            try:
                m = MaxStyle(6, 3, p=0.5, use_gpu=False, **kw)
            except AssertionError:
                entry["raises"] = "AssertionError"
                meta.append(entry)
                idx += 1
                continue
            entry["raises"] = None
            entry["n_params"] = len(list(m.parameters()))
            entry["param_names"] = [k for k, _ in m.named_parameters()]
            entry["state_dict_keys"] = list(m.state_dict().keys())
            # one more CPU draw pins how much of the CPU generator the ctor consumed
            entry["next_cpu_rand"] = float(torch.rand(1).item())
            for k, v in module_state(m).items():
                out[f"c{idx}_{k}"] = v
            meta.append(entry)
            idx += 1
    # small batch sizes: the non-identity redraw loop (maxstyle.py:56-58) triggers often for N=2
    for seed in range(12):
        torch.manual_seed(seed)
        m = MaxStyle(2, 1, p=1.0, use_gpu=False)
        entry = dict(idx=idx, seed=seed, N=2, C=1, p=1.0, mix_style=True, no_noise=False, mix_learnable=True,
                     noise_learnable=True, always_use_beta=False, raises=None, n_params=3,
                     param_names=[k for k, _ in m.named_parameters()],
                     state_dict_keys=list(m.state_dict().keys()),
                     next_cpu_rand=float(torch.rand(1).item()))
        for k, v in module_state(m).items():
            out[f"c{idx}_{k}"] = v
        meta.append(entry)
        idx += 1
    return out, meta


FWD_BWD_CASES = [
    # name, N, C, H, W, input kind, ctor kwargs, lmda override
    ("ragged_m35", 4, 3, 5, 7, "scaled", {}, None),
    ("n2_c1", 2, 1, 8, 8, "scaled", {}, None),
    ("sigmoid_c1", 6, 1, 16, 16, "sigmoid", {}, None),
    ("offset_mean", 5, 4, 12, 12, "offset", {}, None),
    ("mid_16ch", 8, 16, 12, 12, "scaled", {}, None),
    ("wide_plane", 3, 2, 40, 52, "scaled", {}, None),
    ("odd_everything", 7, 5, 9, 11, "scaled", {}, None),
    ("beta_lmda", 6, 4, 10, 10, "scaled", {"always_use_beta": True}, None),
    ("no_mix", 6, 4, 10, 10, "scaled", {"mix_style": False}, None),
    ("no_noise", 6, 4, 10, 10, "scaled", {"no_noise": True, "noise_learnable": False}, None),
    ("fixed_noise", 6, 4, 10, 10, "scaled", {"noise_learnable": False}, None),
    ("fixed_mix", 6, 4, 10, 10, "scaled", {"mix_learnable": False}, None),
    ("lmda_edges", 6, 4, 10, 10, "scaled", {}, [0.0, -0.3, 1.0, 1.7, 0.25, 0.999]),
    ("m2_plane", 4, 3, 1, 2, "scaled", {}, None),
]


def gen_fwd_bwd_cases(MaxStyle):
    out, meta = {}, []
    for i, (name, n, c, h, w, kind, kw, lm) in enumerate(FWD_BWD_CASES):
        seed = 1000 + i
        torch.manual_seed(seed)
        m = MaxStyle(n, c, p=1.0, use_gpu=False, **kw)         # p=1.0: always active
        if lm is not None:
            with torch.no_grad():
                m.lmda.copy_(torch.tensor(lm, dtype=torch.float32).view(n, 1, 1, 1))
        x_np = make_input(seed, (n, c, h, w), kind)
        dy_np = np.random.RandomState(seed + 5000).standard_normal(size=(n, c, h, w)).astype(np.float32)
        x = torch.from_numpy(x_np).clone().requires_grad_(True)
        y = m(x)
        assert y is not x
        y.backward(torch.from_numpy(dy_np))
        # channels_last input gives the same values (NHWC oracle protocol)
        m2_in = torch.from_numpy(x_np).clone().contiguous(memory_format=torch.channels_last)
        y_cl = m(m2_in)
        assert torch.allclose(y_cl, y.detach(), rtol=1e-5, atol=1e-6), name
        mu = x.detach().mean(dim=[2, 3])
        sig = (x.detach().var(dim=[2, 3]) + m.eps).sqrt()
        st = module_state(m)
        pre = f"f{i}_"
        for k, v in st.items():
            out[pre + k] = v
        out[pre + "mu"] = t2n(mu)
        out[pre + "sig"] = t2n(sig)
        out[pre + "gamma_std"] = t2n(m.gamma_std).reshape(c)
        out[pre + "beta_std"] = t2n(m.beta_std).reshape(c)
        out[pre + "y"] = t2n(y)
        out[pre + "dx"] = t2n(x.grad)
        for pname in ("gamma_noise", "beta_noise", "lmda"):
            p_ = getattr(m, pname)
            g = p_.grad if (isinstance(p_, torch.nn.Parameter) and p_.grad is not None) else None
            out[pre + "d_" + pname] = (t2n(g).reshape(n, -1) if g is not None else np.zeros((0,), np.float32))
        meta.append(dict(idx=i, name=name, seed=seed, N=n, C=c, H=h, W=w, kind=kind, kwargs=kw,
                         lmda_override=lm is not None))
    return out, meta


def gen_cache_case(MaxStyle):
    """gamma_std/beta_std are computed on the first forward and reused (maxstyle.py:165-168)
    until reset() (maxstyle.py:136-138)."""
    torch.manual_seed(77)
    n, c, h, w = 5, 3, 6, 6
    m = MaxStyle(n, c, p=1.0, use_gpu=False)
    x1 = make_input(77, (n, c, h, w))
    x2 = make_input(78, (n, c, h, w)) * 3.0
    y1 = m(torch.from_numpy(x1))
    gs1 = t2n(m.gamma_std).reshape(c).copy()
    y2 = m(torch.from_numpy(x2))
    gs2 = t2n(m.gamma_std).reshape(c).copy()
    assert np.array_equal(gs1, gs2)
    out = {f"k_{k}": v for k, v in module_state(m).items()}
    out.update(k_y1=t2n(y1), k_y2=t2n(y2), k_gamma_std=gs1, k_beta_std=t2n(m.beta_std).reshape(c))
    return out, dict(seed=77, seed2=78, scale2=3.0, N=n, C=c, H=h, W=w)


def gen_selftest(MaxStyle):
    """The reference's print-only self test (maxstyle.py:193-241), values recorded."""
    torch.manual_seed(43)
    features = (3 * torch.arange(32, dtype=torch.float32) + 5).view(4, 2, 2, 2)
    m = MaxStyle(batch_size=4, num_feature=2, p=0.5, mix_style=True, mix_learnable=True, noise_learnable=True,
                 always_use_beta=False, no_noise=False, use_gpu=False, debug=False)
    out = {f"s_{k}": v for k, v in module_state(m).items()}
    opt = torch.optim.Adam(list(m.parameters()), lr=0.1)
    loss_fn = torch.nn.MSELoss(reduction="mean")
    gt = torch.ones_like(features)
    losses, ys = [], []
    for i in range(5):
        y = m(features)
        loss = loss_fn(y, gt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
        ys.append(t2n(y))
        for k in ("gamma_noise", "beta_noise", "lmda"):
            out[f"s_step{i}_{k}"] = t2n(getattr(m, k)).reshape(4, -1).copy()
    out["s_losses"] = np.asarray(losses, np.float32)
    out["s_y"] = np.stack(ys)
    out["s_gamma_std"] = t2n(m.gamma_std).reshape(2)
    out["s_beta_std"] = t2n(m.beta_std).reshape(2)
    return out, dict(seed=43)


def gen_loop_case(MaxStyle):
    """5 Adam(lr=0.1) steps on loss = -mean(y*w) (ascent-style, like -CE) on a non-toy shape;
    the trajectory of the three parameters is the multi-step parity vector."""
    seed = 2024
    n, c, h, w = 8, 6, 14, 14
    torch.manual_seed(seed)
    m = MaxStyle(n, c, p=1.0, use_gpu=False)
    x = torch.from_numpy(make_input(seed, (n, c, h, w)))
    wgt = torch.from_numpy(np.random.RandomState(seed + 1).standard_normal(size=(n, c, h, w)).astype(np.float32))
    out = {f"l_{k}": v for k, v in module_state(m).items()}
    opt = torch.optim.Adam(m.parameters(), lr=0.1)
    losses = []
    for i in range(5):
        y = m(x)
        loss = -(torch.tanh(y) * wgt).mean()
        opt.zero_grad()
        loss.backward()
        for k in ("gamma_noise", "beta_noise", "lmda"):
            out[f"l_step{i}_grad_{k}"] = t2n(getattr(m, k).grad).reshape(n, -1).copy()
        opt.step()
        losses.append(loss.item())
        for k in ("gamma_noise", "beta_noise", "lmda"):
            out[f"l_step{i}_{k}"] = t2n(getattr(m, k)).reshape(n, -1).copy()
    out["l_losses"] = np.asarray(losses, np.float32)
    out["l_y_final"] = t2n(m(x))
    return out, dict(seed=seed, N=n, C=c, H=h, W=w)


def main():
    if not os.path.exists(REF_FILE):
        sys.exit(f"reference not found at {REF_FILE}; fixtures can only be generated in the build container")
    torch.set_num_threads(1)          # deterministic reduction order for the recorded values
    MaxStyle = load_reference()
    os.makedirs(OUT_DIR, exist_ok=True)
    manifest = dict(torch=torch.__version__, numpy=np.__version__, reference=REF_FILE,
                    note="generated by oracle/gen_golden.py from the unmodified reference on CPU")
    a, manifest["ctor"] = gen_ctor_cases(MaxStyle)
    np.savez_compressed(os.path.join(OUT_DIR, "ctor.npz"), **a)
    a, manifest["fwd_bwd"] = gen_fwd_bwd_cases(MaxStyle)
    np.savez_compressed(os.path.join(OUT_DIR, "fwd_bwd.npz"), **a)
    a, manifest["cache"] = gen_cache_case(MaxStyle)
    np.savez_compressed(os.path.join(OUT_DIR, "cache.npz"), **a)
    a, manifest["selftest"] = gen_selftest(MaxStyle)
    np.savez_compressed(os.path.join(OUT_DIR, "selftest.npz"), **a)
    a, manifest["loop"] = gen_loop_case(MaxStyle)
    np.savez_compressed(os.path.join(OUT_DIR, "loop.npz"), **a)
    with open(os.path.join(OUT_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", OUT_DIR, {k: len(v) if isinstance(v, list) else 1 for k, v in manifest.items()
                             if k in ("ctor", "fwd_bwd", "cache", "selftest", "loop")})


if __name__ == "__main__":
    main()
