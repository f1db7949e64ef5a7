#!/usr/bin/env python
"""R-rank NCCL parity check of GlobalBatchMaxStyle (run under torchrun; one rank per GPU).

Oracle for an R-rank run = the reference on the concatenated global batch: here the
golden case "mid_16ch" generated from the unmodified reference (tests/golden/fwd_bwd.npz).  Every rank seeds
like the single-device reference, takes its rows, runs forward + backward (+ a fused Adam step) on its GPU,
and compares its slab with the golden tensors.  Tolerances: BASELINE.json's 1e-5 forward, 1e-4 gradients,
perm bit-exact.  Prints one line per rank; exit code != 0 on any mismatch.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from maxstyle_b200 import GlobalBatchMaxStyle, FusedStyleOptimizer
    from oracle import maxstyle_oracle as O
    from oracle.gen_golden import make_input

    man = json.load(open(os.path.join(ROOT, "tests", "golden", "MANIFEST.json")))
    idx = [m["name"] for m in man["fwd_bwd"]].index("mid_16ch")
    meta = man["fwd_bwd"][idx]
    g = np.load(os.path.join(ROOT, "tests", "golden", "fwd_bwd.npz"))
    pre = f"f{idx}_"
    n_glob, c, h, w = meta["N"], meta["C"], meta["H"], meta["W"]
    assert n_glob % world == 0, f"golden batch {n_glob} does not split over {world} ranks"
    n_loc = n_glob // world
    off = rank * n_loc

    torch.manual_seed(meta["seed"])
    layer = GlobalBatchMaxStyle(n_loc, c, p=1.0)
    assert np.array_equal(layer.perm.numpy(), g[pre + "perm"]), "perm differs from the reference's draw"   # bit-exact
    # parameters are drawn on the CUDA generator here (the goldens came from the CPU generator): load the recorded rows
    with torch.no_grad():
        layer.gamma_noise.copy_(torch.from_numpy(g[pre + "gamma_noise"][off:off + n_loc]).view(n_loc, c, 1, 1))
        layer.beta_noise.copy_(torch.from_numpy(g[pre + "beta_noise"][off:off + n_loc]).view(n_loc, c, 1, 1))
        layer.lmda.copy_(torch.from_numpy(g[pre + "lmda"].reshape(-1)[off:off + n_loc]).view(n_loc, 1, 1, 1))
    x_np = make_input(meta["seed"], (n_glob, c, h, w), meta["kind"])
    dy_np = np.random.RandomState(meta["seed"] + 5000).standard_normal(size=(n_glob, c, h, w)).astype(np.float32)

    def close(a, b, rtol, name, scale=None):
        a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
        s = np.abs(b).max() if scale is None else scale
        err = np.abs(a - b).max()
        assert err <= rtol * max(s, 1e-30), f"rank {rank} {name}: err {err:.3e} scale {s:.3e} rtol {rtol:g}"
        return err / max(s, 1e-30)

    errs = {}
    for fmt, tag in ((torch.contiguous_format, "nchw"), (torch.channels_last, "nhwc")):
        layer.gamma_std = layer.beta_std = None
        layer.zero_grad()
        x = torch.from_numpy(x_np[off:off + n_loc]).to(dev).contiguous(memory_format=fmt).requires_grad_(True)
        dy = torch.from_numpy(dy_np[off:off + n_loc]).to(dev).contiguous(memory_format=fmt)
        y = layer(x)
        y.backward(dy)
        t2n = lambda t: t.detach().float().cpu().numpy()
        errs[tag] = dict(
            y=close(t2n(y), g[pre + "y"][off:off + n_loc], 1e-5, "y"),
            dx=close(t2n(x.grad), g[pre + "dx"][off:off + n_loc], 1e-4, "dx"),
            d_gamma=close(t2n(layer.gamma_noise.grad).reshape(n_loc, c), g[pre + "d_gamma_noise"][off:off + n_loc], 1e-4, "d_gamma",
                          scale=np.abs(g[pre + "d_gamma_noise"]).max()),
            d_beta=close(t2n(layer.beta_noise.grad).reshape(n_loc, c), g[pre + "d_beta_noise"][off:off + n_loc], 1e-4, "d_beta",
                         scale=np.abs(g[pre + "d_beta_noise"]).max()),
            d_lmda=close(t2n(layer.lmda.grad).reshape(-1), g[pre + "d_lmda"].reshape(-1)[off:off + n_loc], 1e-4, "d_lmda",
                         scale=np.abs(g[pre + "d_lmda"]).max()),
            gamma_std=close(t2n(layer.gamma_std).reshape(-1), g[pre + "gamma_std"].reshape(-1), 1e-5, "gamma_std",
                            scale=np.abs(g[pre + "sig"]).max()),
        )
    # fused Adam step on the local rows == the oracle's restatement of torch.optim.Adam on those rows
    before = layer.lmda.detach().cpu().numpy().reshape(-1).copy()
    grad = layer.lmda.grad.detach().cpu().numpy().reshape(-1).copy()
    opt = FusedStyleOptimizer([layer], lr=0.1)
    layer.zero_grad()
    layer(x).backward(dy)
    want = O.adam_step(before, grad, O.AdamState(np.zeros(n_loc, np.float32), np.zeros(n_loc, np.float32)))
    errs["adam_lmda"] = float(np.abs(layer.lmda.detach().cpu().numpy().reshape(-1) - want).max())
    assert errs["adam_lmda"] < 1e-5, errs
    # GraphedLayerStep on the global-batch layer, with both transports of the (mu | sig) rows: the fused peer-memory
    # exchange + tables kernel (maxstyle_tables_p2p) and NCCL's all-gather captured in the graph.  Several replays walk
    # the exchange's epoch / parity scheme; every one must reproduce the golden slab.
    from maxstyle_b200 import GraphedLayerStep
    from maxstyle_b200 import functional as F_
    layer._fused_step = None                                   # gradients only: the state stays fixed across replays
    with torch.no_grad():                                      # the fused Adam step above moved all three
        layer.gamma_noise.copy_(torch.from_numpy(g[pre + "gamma_noise"][off:off + n_loc]).view(n_loc, c, 1, 1))
        layer.beta_noise.copy_(torch.from_numpy(g[pre + "beta_noise"][off:off + n_loc]).view(n_loc, c, 1, 1))
        layer.lmda.copy_(torch.from_numpy(g[pre + "lmda"].reshape(-1)[off:off + n_loc]).view(n_loc, 1, 1, 1))
    layer.gamma_std = layer.beta_std = None
    xs = torch.from_numpy(x_np[off:off + n_loc]).to(dev)
    dys = torch.from_numpy(dy_np[off:off + n_loc]).to(dev)
    for transport in ("p2p-one-kernel", "p2p", "nccl"):
        layer.gamma_std = layer.beta_std = None                 # each variant also runs the first-forward (batch std) path
        try:
            gs = GraphedLayerStep(layer, xs, dys, exchange=transport.split("-")[0], one_kernel=transport == "p2p-one-kernel")
        except RuntimeError as e:
            if "peer-memory exchange requested but not available" not in str(e):
                raise
            errs[f"graph_{transport}"] = "symmetric memory unavailable on this system: skipped (every rank agrees)"
            continue
        assert gs.exchange == transport.split("-")[0]
        errs[f"graph_{transport}_gamma_std"] = close(t2n(layer.gamma_std).reshape(-1), g[pre + "gamma_std"].reshape(-1), 1e-5,
                                                     f"graphed {transport} gamma_std", scale=np.abs(g[pre + "sig"]).max())
        for rep in range(5):
            y_g, dx_g = gs.run()
            torch.cuda.synchronize()
            errs[f"graph_{transport}_y"] = close(t2n(y_g), g[pre + "y"][off:off + n_loc], 1e-5, f"graphed {transport} y (replay {rep})")
            errs[f"graph_{transport}_dx"] = close(t2n(dx_g), g[pre + "dx"][off:off + n_loc], 1e-4, f"graphed {transport} dx (replay {rep})")
            errs[f"graph_{transport}_d_lmda"] = close(t2n(gs.grads[2]), g[pre + "d_lmda"].reshape(-1)[off:off + n_loc], 1e-4,
                                                      f"graphed {transport} d_lmda", scale=np.abs(g[pre + "d_lmda"]).max())
        if transport.startswith("p2p"):
            gs.peer.check()
            assert int(gs.peer.epoch.item()) == 7, int(gs.peer.epoch.item())     # first forward + warm-up + 5 replays
            F_.workspace_status(gs.ws, n_loc, c, h, w, 0)                        # no device-side wait timed out
        errs[f"graph_{transport}_kernels"] = float(gs.kernels_per_step)
        gs.close()
        torch.cuda.synchronize()
        dist.barrier()
    # The one-kernel multi-GPU forward (maxstyle_fwd_p2p: L2-window kernel, exchange in the channel finaliser) needs planes of
    # >= 64 KB, which the golden case does not have: on a synthetic 100 KB-plane problem it must agree with the NCCL transport
    # (different statistics kernels: last-bit differences only), first forward (global batch std) and cached forwards alike.
    torch.manual_seed(77)
    big = GlobalBatchMaxStyle(6, 8, p=1.0)
    gen = torch.Generator(device=dev).manual_seed(500 + rank)
    xb = torch.randn(6, 8, 160, 160, device=dev, generator=gen) * 1.7 + 0.4
    dyb = torch.randn(6, 8, 160, 160, device=dev, generator=gen)
    outs = {}
    for name, kw in (("one", dict(exchange="p2p", one_kernel=True)), ("nccl", dict(exchange="nccl"))):
        big.gamma_std = big.beta_std = None
        try:
            gsb = GraphedLayerStep(big, xb, dyb, **kw)
        except RuntimeError as e:
            if "peer-memory exchange requested but not available" not in str(e):
                raise
            continue
        for _ in range(3):
            yb, dxb = gsb.run()
        torch.cuda.synchronize()
        outs[name] = [t.detach().clone() for t in (yb, dxb, gsb.grads[0], gsb.grads[1], gsb.grads[2], big.gamma_std, big.beta_std)]
        if name == "one":
            assert gsb.one_kernel is True and gsb.kernels_per_step in (2, 3), (gsb.one_kernel, gsb.kernels_per_step)   # [barrier +] forward + backward
            gsb.peer.check()
            F_.workspace_status(gsb.ws, 6, 8, 160, 160, 0)
        gsb.close()
        torch.cuda.synchronize()
        dist.barrier()
    for i, nm in enumerate(("y", "dx", "d_gamma", "d_beta", "d_lmda", "gamma_std", "beta_std") if "one" in outs else ()):
        a_, b_ = t2n(outs["one"][i]), t2n(outs["nccl"][i])
        errs[f"one_kernel_vs_nccl_{nm}"] = close(a_, b_, 1e-5 if nm in ("y", "gamma_std", "beta_std") else 1e-4, f"one-kernel vs nccl {nm}")
    # ---- the path bench.py / SCALE time: per-rank planes of 196 KB (CTA-mode statistics / apply / backward, the one-kernel
    # paired forward with its peer pushes, maxstyle_tables_p2p at world x 8 rows), every transport, first forward + replays,
    # against the float64 oracle on the CONCATENATED batch (bench.parity_check: reference semantics, maxstyle.py:157-185 with
    # perm over the global batch) -- y <= 1e-5, gradients <= 1e-4, batch std <= 1e-5, perm bit-exact.
    from bench import parity_check
    seed = 4321
    nb, cb, hb, wb = 8, 8, 224, 224
    for transport in ("p2p-one-kernel", "p2p", "nccl"):
        torch.manual_seed(seed)
        lay = GlobalBatchMaxStyle(nb, cb, p=1.0)
        genb = torch.Generator(device=dev).manual_seed(900 + rank)
        xr = torch.randn(nb, cb, hb, wb, device=dev, generator=genb) * (1.0 + 0.3 * rank) + 0.2 * rank
        dyr = torch.randn(nb, cb, hb, wb, device=dev, generator=genb)
        try:
            gsr = GraphedLayerStep(lay, xr, dyr, exchange=transport.split("-")[0], one_kernel=transport == "p2p-one-kernel")
        except RuntimeError as e:
            if "peer-memory exchange requested but not available" not in str(e):
                raise
            errs[f"bench_path_{transport}"] = "symmetric memory unavailable on this system: skipped (every rank agrees)"
            continue
        if transport == "p2p-one-kernel":
            assert gsr.one_kernel is True and gsr.kernels_per_step == 2, (gsr.one_kernel, gsr.kernels_per_step)
        for _ in range(4):
            gsr.run()
        par = parity_check(lay, gsr, world, rank, dev, seed)
        assert par["ok"], f"rank {rank} {transport}: {par}"
        errs[f"bench_path_{transport}"] = {k: par[k] for k in ("y", "dx", "d_gamma", "d_beta", "d_lmda", "gamma_std", "beta_std")}
        gsr.close()
        torch.cuda.synchronize()
        dist.barrier()
    # ---- config-5 planes (512 x 512 = 1 MiB, 16 pieces each): a channel's N * P pieces exceed the grid, so the one-kernel forward
    # takes the samples in cycle order of the GLOBAL permutation (pair_fwd.cuh) -- partners one step ahead, on whichever rank.
    torch.manual_seed(seed + 1)
    lay = GlobalBatchMaxStyle(40, 2, p=1.0)
    genc = torch.Generator(device=dev).manual_seed(1900 + rank)
    xr = torch.randn(40, 2, 512, 512, device=dev, generator=genc) * (1.0 + 0.2 * rank) - 0.1 * rank
    dyr = torch.randn(40, 2, 512, 512, device=dev, generator=genc)
    try:
        gsr = GraphedLayerStep(lay, xr, dyr, exchange="p2p", one_kernel=True)
    except RuntimeError as e:
        if "peer-memory exchange requested but not available" not in str(e):
            raise
        gsr = None
        errs["global_cycle_order"] = "symmetric memory unavailable on this system: skipped (every rank agrees)"
    if gsr is not None:
        assert gsr.one_kernel is True and gsr.kernels_per_step == 2, (gsr.one_kernel, gsr.kernels_per_step)
        for _ in range(4):
            gsr.run()
        par = parity_check(lay, gsr, world, rank, dev, seed + 1)
        assert par["ok"], f"rank {rank} global cycle order: {par}"
        errs["global_cycle_order"] = {k: par[k] for k in ("y", "dx", "d_gamma", "d_beta", "d_lmda", "gamma_std", "beta_std")}
        gsr.close()
        torch.cuda.synchronize()
        dist.barrier()
    print(f"[dist_parity] rank {rank}/{world} ok", json.dumps({k: (v if isinstance(v, (float, str)) else {a: f"{b:.1e}" for a, b in v.items()})
                                                                for k, v in errs.items()}), flush=True)
    sys.stdout.flush()
    os._exit(0)        # graphs with captured NCCL kernels were just released; skip the communicator teardown


if __name__ == "__main__":
    main()
