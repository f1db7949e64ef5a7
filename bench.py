#!/usr/bin/env python
"""Benchmark of the MaxStyle layer hot path (BASELINE.json metric: fwd+bwd samples/s and
achieved HBM GB/s on B200 vs the host CPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic feature maps: forward,
backward (dX + the three parameter gradients) and the optimiser step fused into the backward
epilogue.  Workload = BASELINE config 1 shape (the FCN_64 decoder's layer-4 activation of
config 2): fp32 NCHW, 20 x 64 x 224 x 224 per GPU, x ~ N(0,1)*1.5+0.25, dy ~ N(0,1).  With N > 1
the batch shards data-parallel (20 per GPU, weak scaling) and mixing partners come from the
global batch through the one all-gather of the [N,2C] (mu|sig) tables.

Prints ONE JSON line (rank 0).  `value` times the device-resident path with CUDA events,
`e2e` the same step through the public module API from pinned HOST buffers (H2D of x and dy,
D2H of y, dX and the parameter gradients inside the timed region).  `roofline` is the dominant
kernel (the backward sweep) against the measured HBM copy peak of MEASURED_PEAKS.json;
`cpu_baseline` is the reference's op chain (oracle/torch_port.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "maxstyle_layer_fwd_bwd_step_samples_per_sec"
UNIT = "samples/s"
SHAPE = dict(N=20, C=64, H=224, W=224)            # per GPU
HBM_FALLBACK_GBS = 6650.0                         # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def workload_config(n_gpus: int) -> dict:
    s = SHAPE
    return {
        "workload": f"MaxStyle layer fwd+bwd+fused Adam step, fp32 NCHW {s['N']}x{s['C']}x{s['H']}x{s['W']} per GPU "
                    "(BASELINE config-1 shape = FCN_64 layer-4 activation of config 2), all params learnable, p=1",
        "per_gpu_batch": s["N"], "global_batch": s["N"] * n_gpus, "channels": s["C"], "height": s["H"], "width": s["W"],
        "layout": "NCHW", "parallelism": f"dp{n_gpus}" + ("+exchange(mu,sig)" if n_gpus > 1 else ""),
        "l2_policy": "working set 1.03 GB per GPU (x, y, dy, dx of 257 MB each) >> 126 MB L2; no explicit flush",
        "execution": ("eager module path" if os.environ.get("BENCH_EAGER", "0") == "1" else
                      "GraphedLayerStep: two CUDA graphs per step (forward | backward + fused Adam), replayed"),
    }


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("bwd_nchw_kernel", {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled through NVML from a background thread while the timed region
    runs (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; a polling
    nvidia-smi subprocess stalls kernel launches while it starts up, NVML calls from this process do not)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int, period_s: float = 0.02):
        self.gpu, self.period = gpu_index, period_s
        self.samples, self.thread, self.stop_flag, self.err = [], None, False, None

    def start(self):
        import threading
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if visible:
                try:
                    idx = int(visible.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception as e:           # noqa: BLE001
            self.err = f"nvml unavailable: {e}"
            return

        def loop():
            while not self.stop_flag:
                try:
                    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                    self.samples.append((time.perf_counter(), sm, rs, pw))
                except Exception as e:   # noqa: BLE001
                    self.err = str(e)
                    return
                time.sleep(self.period)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None) -> dict:
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        inside = [s for s in self.samples if (t0 is None or s[0] >= t0) and (t1 is None or s[0] <= t1)] or self.samples
        reasons = sorted(k for k, bit in self.REASONS.items() if any(s[2] & bit for s in inside))
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": float(self.max_sm),
                "power_w_max": max(s[3] for s in inside), "samples": len(inside), "reasons": reasons,
                "source": "NVML (clocks.sm / clocks_event_reasons), sampled every 20 ms from the start of the timed region "
                          "to the end of the load-hold loop that follows it"}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port) on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_timing(n, c, h, w, steps, warmup, threads, budget_s=None):
    """The reference's CPU implementation of the path on the host cores: the UNMODIFIED reference layer from oracle/_ref
    (kind "reference") when it has been staged, else the validated port of its op chain (kind "port")."""
    try:
        from oracle import ref_shims
        if ref_shims.reference_root() is not None:
            from oracle.ref_layer_bench import time_reference_layer
            res = time_reference_layer(n, c, h, w, steps=steps, warmup=warmup, threads=threads, budget_s=budget_s)
            res["what"] = ("unmodified reference MaxStyle(use_gpu=False) (src/advanced/maxstyle.py, staged in oracle/_ref): "
                           "zero_grad, forward, backward, torch.optim.Adam(lr=0.1).step()")
            return res
    except Exception as e:           # noqa: BLE001  (fall back to the port, and say so)
        print(f"[bench] reference layer unavailable ({e!r}); timing the port", file=sys.stderr)
    from oracle.torch_port import time_cpu_baseline
    res = time_cpu_baseline(n, c, h, w, budget_s=budget_s or 60.0, min_iters=steps, warmup=warmup, threads=threads)
    res["kind"] = "port"
    res["what"] = "port of the reference's ATen op chain + autograd + torch.optim.Adam (oracle/torch_port.py)"
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    s = SHAPE
    # a "step" is the full per-GPU batch of the workload; exactly --steps timed steps after --warmup untimed ones, unless the
    # time budget (a few minutes at most) runs out first -- `steps` in the line says how many were timed
    budget = float(os.environ.get("BENCH_REF_BUDGET_S", "150"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    res = cpu_reference_timing(s["N"], s["C"], s["H"], s["W"], steps, warmup, threads, budget_s=budget)
    value = res["samples_per_s"]
    sample = (f"{res['iters']} timed steps (mean) of the full per-GPU batch {s['N']}x{s['C']}x{s['H']}x{s['W']} fp32 after {warmup} warm-up: "
              + res["what"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": res["iters"],
        "warmup": warmup, "ms_per_step": res["seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["threads"], "kind": res["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference's own CPU implementation of the path on the box's host cores, one process (rank 0), all threads",
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# untimed checks that ride along with the benchmark
# --------------------------------------------------------------------------------------------
def parity_check(layer, gstep, world, rank, dev, seed):
    """This rank's outputs on the benchmarked buffers against the float64 oracle (oracle/maxstyle_oracle.py, pinned to the
    reference by tests/test_oracle_golden.py), run AFTER the timed region: y through the very forward graph that was timed,
    dX and the parameter gradients through the same backward kernel (without the fused step, so that the parameters stay the
    ones the oracle is given).  Multi GPU: the oracle is the reference semantics on the CONCATENATED batch
    (maxstyle.py:157-185 with perm over the global batch): every rank takes the float64 statistics of its own rows, the tables
    are all-gathered (test plumbing, not the product path) and each rank checks its slice.  Errors are max-norm relative
    (max|a-b| / max|b|), maximum over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from maxstyle_b200 import functional as F
    from oracle import maxstyle_oracle as O
    n, c, h, w = gstep.x.shape
    f64 = np.float64
    # After hundreds of Adam(lr=0.1) steps on random gradients lmda has left [0, 1] on most samples, where the clamp's gradient
    # is zero (the reference's behaviour) -- d_lmda would be checked as 0 == 0.  Put it back inside.
    with torch.no_grad():
        gen = torch.Generator().manual_seed(4242 + rank)
        layer.lmda.copy_((torch.rand(n, generator=gen) * 0.9 + 0.05).view_as(layer.lmda))
    gam = layer.gamma_noise.detach().float().cpu().numpy().reshape(n, c).astype(f64)
    bet = layer.beta_noise.detach().float().cpu().numpy().reshape(n, c).astype(f64)
    lm = layer.lmda.detach().float().cpu().numpy().reshape(n).astype(f64)
    gs = layer.gamma_std.detach().float().cpu().numpy().reshape(c).astype(f64)
    bs = layer.beta_std.detach().float().cpu().numpy().reshape(c).astype(f64)
    y = gstep.forward().clone()
    dx, dg, db, dl = F.backward_raw(gstep.dy, gstep.x, gstep.mu_all, gstep.sig_all, gstep.row_offset, gstep.scale, gstep.perm,
                                    layer.lmda, layer.gamma_std, layer.beta_std, gstep.flags, gstep.ws, need_dx=True)
    torch.cuda.synchronize()
    F.workspace_status(gstep.ws, n, c, h, w, F.dtype_code(gstep.x), F.layout_of(gstep.x))     # raises if a device-side wait gave up
    if getattr(gstep, "peer", None) is not None:
        gstep.peer.check()
    # Tensors too large for a float64 numpy pass in seconds (config 5: 2 G elements) are checked on the first `k` samples of
    # this rank -- every channel of them, so that d_lmda (a sum over channels) is covered; the statistics of all other rows are
    # then the exchanged table the kernels produced (each rank checks its own rows of it).
    full = gstep.x.numel() <= (1 << 27)
    k = n if full else min(n, 4)
    x_np, dy_np = gstep.x[:k].cpu().numpy(), gstep.dy[:k].cpu().numpy()
    y, dx, dg, db, dl = y[:k], dx[:k], dg[:k], db[:k], dl[:k]
    gam, bet, lm = gam[:k], bet[:k], lm[:k]
    mu, sig = O.instance_stats(x_np, layer.eps, f64)
    row_offset = gstep.row_offset
    if full and world > 1:
        mine = torch.from_numpy(np.concatenate([mu, sig], axis=1)).to(dev)
        table = torch.empty(world * n, 2 * c, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(table, mine)
        table = table.cpu().numpy()
        g_mu, g_sig = table[:, :c], table[:, c:]
    elif full:
        g_mu, g_sig = mu, sig
    else:
        g_mu = gstep.mu_all.detach().double().cpu().numpy().copy()
        g_sig = gstep.sig_all.detach().double().cpu().numpy().copy()
        g_mu[row_offset:row_offset + k] = mu
        g_sig[row_offset:row_offset + k] = sig
    use_global = world > 1 or not full
    st = O.StyleState(perm=layer.perm.numpy(), gamma_noise=gam, beta_noise=bet, lmda=lm, p=1.0, gamma_std=gs, beta_std=bs)
    y64, cache = O.forward(x_np, st, f64, global_mu=g_mu if use_global else None, global_sig=g_sig if use_global else None,
                           row_offset=row_offset)
    dx64, dg64, db64, dl64 = O.backward(dy_np, x_np, st, cache, f64, global_mu=g_mu if use_global else None,
                                        global_sig=g_sig if use_global else None, row_offset=row_offset)

    def rel(a, b):
        a = a.detach().double().cpu().numpy().reshape(b.shape)
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

    # the reference's draw of the permutation under this seed (maxstyle.py:55-58: CPU randperm, redrawn while it is the identity)
    torch.manual_seed(seed)
    ng = world * n
    want = torch.randperm(ng)
    while torch.equal(want, torch.arange(ng)):
        want = torch.randperm(ng)
    perm_exact = bool(torch.equal(want, layer.perm) and torch.equal(gstep.perm.cpu(), layer.perm.to(torch.int64)))
    errs = {"y": rel(y, y64), "dx": rel(dx, dx64), "d_gamma": rel(dg, dg64), "d_beta": rel(db, db64), "d_lmda": rel(dl, dl64),
            "gamma_std": float(np.abs(gs - O.batch_std(g_sig, f64)).max() / np.abs(gs).max()),
            "beta_std": float(np.abs(bs - O.batch_std(g_mu, f64)).max() / np.abs(bs).max())}
    t = torch.tensor(list(errs.values()) + [0.0 if perm_exact else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    vals = t.tolist()
    out = {k: vals[i] for i, k in enumerate(errs)}
    out["perm_exact"] = vals[-1] == 0.0
    out["d_lmda_max_abs"] = float(np.abs(dl64).max())        # nonzero: the mixing-weight gradient was really exercised
    out["ok"] = bool(out["y"] <= 1e-5 and out["perm_exact"] and all(out[k] <= 1e-4 for k in ("dx", "d_gamma", "d_beta", "d_lmda"))
                     and out["gamma_std"] <= 1e-4 and out["beta_std"] <= 1e-4)
    # gamma_std / beta_std are standard deviations over the batch of per-plane sigmas / means that differ only in their 3rd-4th
    # digit when planes are large (condition number mean / std ~ 1e2..1e3): fp32 tables cannot hold them to 1e-5 relative, the
    # reference's own fp32 value is as far from the float64 one -- they are held to the gradient tolerance.
    out["tolerance"] = {"y": 1e-5, "gradients": 1e-4, "batch_std": 1e-4, "norm": "max|a-b| / max|b| over the tensor, max over ranks"}
    out["oracle"] = ("float64 numpy oracle on the concatenated global batch, computed on this box after the timed region; y from the "
                     "timed forward graph, gradients from the timed backward kernel (fused step off)"
                     + ("" if full else f"; sampled: the first {k} samples of every rank, all channels"))
    return out


def link_ceiling(dev, world, h2d_dst, d2h_src, iters=4):
    """Bare pinned-memory copies, one H2D and one D2H stream at once on every rank at the same time: what the host link gives
    this rank when all N ranks copy together -- the ceiling `e2e` can reach.  GB/s each way (the slower direction)."""
    import torch
    import torch.distributed as dist
    nbytes = h2d_dst.numel() * h2d_dst.element_size()
    hin = torch.empty(h2d_dst.shape, dtype=h2d_dst.dtype, pin_memory=True)
    hout = torch.empty(d2h_src.shape, dtype=d2h_src.dtype, pin_memory=True)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(k):
        for _ in range(k):
            with torch.cuda.stream(s1):
                h2d_dst.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(d2h_src, non_blocking=True)
        s1.synchronize(); s2.synchronize()

    run(1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(iters)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return nbytes * iters / float(t.item()) / 1e9


# --------------------------------------------------------------------------------------------
# BASELINE.json configs 3 and 5 (`--config`): the other sharded workloads north_star names, same step, same checks
# --------------------------------------------------------------------------------------------
CONFIG_LAYERS = {
    # config 3: Prostate-shaped 192 x 192, batch 32 per GPU, the FCN_16 decoder's three splice points
    3: dict(name="config 3: Prostate-shaped 192x192, batch 32 per GPU, MaxStyle after the last 3 decoder blocks (FCN_16 widths)",
            layers=[(32, 16, 96, 96), (32, 16, 192, 192), (32, 1, 192, 192)]),
    # config 5: large-batch data-parallel step, N = 256 per GPU, C = 32, 512 x 512 (1 MiB planes, 8.6 GB per tensor)
    5: dict(name="config 5: N=256 per GPU, C=32, 512x512 fp32", layers=[(256, 32, 512, 512)]),
}


def run_config(args, cfg_id):
    """One JSON line for config 3 or 5: K timed steps of forward + backward + fused Adam step over every layer of the config
    (GraphedLayerStep per layer), CUDA events, max over ranks, then the untimed oracle check of every layer.  No host-copy leg."""
    import torch
    import torch.distributed as dist
    from maxstyle_b200 import MaxStyle, FusedStyleOptimizer, GlobalBatchMaxStyle, GraphedLayerStep
    from maxstyle_b200 import functional as F
    cfg = CONFIG_LAYERS[cfg_id]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup, K = max(3, args.warmup), max(1, args.steps)
    seed = 1234
    layers, steps = [], []
    for li, (n, c, h, w) in enumerate(cfg["layers"]):
        torch.manual_seed(seed + li)
        layer = GlobalBatchMaxStyle(n, c, p=1.0) if world > 1 else MaxStyle(n, c, p=1.0)
        FusedStyleOptimizer([layer], lr=0.1, mode="adam")
        gen = torch.Generator(device=dev).manual_seed(100 + rank + 17 * li)
        x = torch.randn(n, c, h, w, device=dev, generator=gen) * 1.5 + 0.25
        dy = torch.randn(n, c, h, w, device=dev, generator=gen)
        layers.append(layer)
        steps.append(GraphedLayerStep(layer, x, dy, need_dx=(li > 0 or len(cfg["layers"]) == 1),
                                      exchange=os.environ.get("BENCH_EXCHANGE", "auto")))

    def step():
        for g in steps:
            g.forward()
        for g in reversed(steps):
            g.backward()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = F.launches.kernels
    barrier()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    launches = F.launches.kernels - launches0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    parity = None
    if not args.no_parity:
        per = [parity_check(l, g, world, rank, dev, seed + li) for li, (l, g) in enumerate(zip(layers, steps))]
        parity = {k: max(p[k] for p in per) for k in ("y", "dx", "d_gamma", "d_beta", "d_lmda", "gamma_std", "beta_std")}
        parity["perm_exact"] = all(p["perm_exact"] for p in per)
        parity["ok"] = all(p["ok"] for p in per)
        parity["tolerance"] = per[0]["tolerance"]
    if rank == 0:
        peak, peak_src = hbm_peak()
        n0 = cfg["layers"][0][0]
        bytes_step = sum((5 if (li > 0 or len(cfg["layers"]) == 1) else 4) * n * c * h * w * 4 for li, (n, c, h, w) in enumerate(cfg["layers"]))
        ms = total_ms / K
        line = {"metric": METRIC, "value": world * n0 * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["name"] + "; fwd+bwd+fused Adam step of every layer, all params learnable, p=1",
                           "layers_per_gpu": [list(s) for s in cfg["layers"]], "global_batch": n0 * world,
                           "parallelism": f"dp{world}" + ("+exchange(mu,sig)" if world > 1 else ""),
                           "exchange": [g.exchange for g in steps] if world > 1 else None,
                           "one_kernel_forward": [bool(g.one_kernel) if world > 1 else (g.fwd_kernels == 1) for g in steps],
                           "note": "the first spliced layer needs no dX (its input is detached, encoder_decoder.py:602): 4*E*s there, 5*E*s elsewhere"},
                "step_roofline": {"algorithmic_bytes_per_step": bytes_step, "achieved_GBps": bytes_step / (ms * 1e-3) / 1e9,
                                  "frac_of_peak": bytes_step / (ms * 1e-3) / 1e9 / peak, "peak": peak, "peak_source": peak_src},
                "gpu_launches": launches, "parity": parity}
        print(json.dumps(line), flush=True)
    if world > 1:
        import gc
        import threading
        barrier()
        for g in steps:
            g.close()
        steps.clear()
        gc.collect()
        torch.cuda.synchronize()
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(10.0)
        sys.stdout.flush()
        os._exit(0)
    return 0


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle check of this rank's outputs")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3, 5],
                    help="BASELINE.json config: 1 (default, the metric's config), 3 (192x192 batch 32, three layers) or 5 (256x32x512x512 per GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != 1:
        return run_config(args, args.config)

    import torch
    import torch.distributed as dist
    from maxstyle_b200 import MaxStyle, FusedStyleOptimizer, GlobalBatchMaxStyle, GraphedLayerStep
    from maxstyle_b200 import functional as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    K = max(1, args.steps)

    s = SHAPE
    n, c, h, w = s["N"], s["C"], s["H"], s["W"]
    E = n * c * h * w
    torch.manual_seed(1234)                                  # identical on every rank: global-batch state contract
    layer = (GlobalBatchMaxStyle(n, c, p=1.0) if world > 1 else MaxStyle(n, c, p=1.0))
    opt = FusedStyleOptimizer([layer], lr=0.1, mode="adam")
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    x = (torch.randn(n, c, h, w, device=dev, generator=gen) * 1.5 + 0.25).requires_grad_(True)
    dy = torch.randn(n, c, h, w, device=dev, generator=gen)

    # The step is issued through the package's graphed entry point (maxstyle_b200.GraphedLayerStep): the same C-ABI calls
    # the module's forward / backward make, captured once into two CUDA graphs (forward | backward + fused step) over the
    # resident x / dy and replayed -- the eager module path costs 190-330 us of host work per step, more than the kernels.
    # BENCH_EAGER=1 times the eager module path (layer(x); y.backward(dy); opt.step()) instead.
    eager = os.environ.get("BENCH_EAGER", "0") == "1"
    gstep = None if eager else GraphedLayerStep(layer, x.detach(), dy, exchange=os.environ.get("BENCH_EXCHANGE", "auto"))
    gstep_exchange = ({"p2p": ("one-kernel forward, (mu|sig) rows exchanged over NVLink peer memory in its channel finaliser (maxstyle_fwd_p2p)"
                               if gstep is not None and gstep.one_kernel else
                               "fused exchange+tables kernel over NVLink peer memory (maxstyle_tables_p2p)"),
                       "nccl": "NCCL all_gather_into_tensor captured in the forward graph"}.get(gstep.exchange, "NCCL all-gather (eager module)")
                      if gstep is not None else "NCCL all-gather (eager module)")

    def fwd():
        if eager:
            return layer(x)
        return gstep.forward()

    def bwd(y):
        if eager:
            x.grad = None
            y.backward(dy)
            opt.step()
        else:
            gstep.backward()

    def step():
        bwd(fwd())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_region0 = time.perf_counter()
    # ---- timed region: exactly K steps, CUDA events on the launching (current) stream ------
    # Two events per step bracket the dominant kernel (the backward sweep) for the roofline line.
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = F.launches.kernels
    barrier()
    e_begin.record()
    for i in range(K):
        y = fwd()
        ev[i][0].record()
        bwd(y)
        ev[i][1].record()
    e_end.record()
    barrier()
    launches = F.launches.kernels - launches0
    total_ms = e_begin.elapsed_time(e_end)
    bwd_ms = statistics.fmean(ev[i][0].elapsed_time(ev[i][1]) for i in range(K))
    fwd_ms = total_ms / K - bwd_ms
    # keep the same load running (untimed) long enough for nvidia-smi's 100 ms sampling to see it
    t_hold = time.perf_counter()
    while rank == 0 and world == 1 and time.perf_counter() - t_hold < 1.0:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    if world > 1:
        for _ in range(200):
            step()
        barrier()
    clocks = sampler.stop(t_region0, time.perf_counter()) if rank == 0 else None

    t = torch.tensor([total_ms, fwd_ms, bwd_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, fwd_ms, bwd_ms = (float(v) for v in t.tolist())
    ms_per_step = total_ms / K
    value = world * n * K / (total_ms * 1e-3)

    # ---- untimed: this rank's outputs against the oracle, on the path that was just timed -------------
    parity = None
    if not args.no_parity and gstep is not None:
        parity = parity_check(layer, gstep, world, rank, dev, 1234)

    # ---- e2e: same step through the public module API from pinned host buffers --------------
    e2e = None
    if not args.no_e2e:
        # The public host-buffer call: HostStepPipeline.submit(hx, hdy) -> wait(ticket).  Every step copies THAT
        # step's x and dy from pinned host memory and its y, dX and updated parameters back; the link is full duplex,
        # so the copy-in of step i+1 is enqueued while the copy-out of step i drains (two slots in flight).
        from maxstyle_b200 import HostStepPipeline
        hx = torch.empty(n, c, h, w, dtype=torch.float32, pin_memory=True).copy_(x.detach())
        hdy = torch.empty(n, c, h, w, dtype=torch.float32, pin_memory=True).copy_(dy)
        pipe = HostStepPipeline(layer, (n, c, h, w), torch.float32, depth=2)
        check = 0.0

        def e2e_run(k):
            nonlocal check
            tickets = []
            for i in range(k):
                tickets.append(pipe.submit(hx, hdy))
                if i >= 1:
                    res = pipe.wait(tickets[i - 1])          # the caller reads step i-1's results on the host
                    check += float(res.params[0])
            res = pipe.wait(tickets[-1])
            check += float(res.params[0])

        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        ke = max(2, args.e2e_steps)
        e2e_run(ke)
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        del pipe, hx, hdy
        ceiling = link_ceiling(dev, world, gstep.y if gstep is not None else torch.empty_like(dy), dy)
        e2e = {"value": world * n * ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * E * 4,
               "d2h_bytes_per_step": 2 * E * 4 + (2 * n * c + n) * 4, "steps": ke, "ms_per_step": e2e_s / ke * 1e3,
               "link_GBps_each_way": 2 * E * 4 / (e2e_s / ke) / 1e9,
               "link_ceiling_GBps_each_way": ceiling, "frac_of_link_ceiling": 2 * E * 4 / (e2e_s / ke) / 1e9 / ceiling,
               "link_ceiling_how": "bare pinned copies, one H2D and one D2H stream at once, all ranks at the same time (257 MB each way per copy)",
               "note": "HostStepPipeline (public API): pinned host x,dy -> H2D -> layer fwd/bwd/fused step -> D2H y,dX,params "
                       "every step; 2 steps in flight so copy-in and copy-out share the full-duplex link; PCIe-bound"}

    if rank == 0:
        peak, peak_src = hbm_peak()
        bwd_bytes = 3 * E * 4                                  # read dy, read x, write dX
        step_bytes = 5 * E * 4                                 # + forward: read x, write y
        achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), **({"exchange": gstep_exchange} if world > 1 else {})),
            "roofline": {"bound": "hbm", "kernel": "bwd_nchw_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bwd_bytes, "avg_launch_ms": bwd_ms},
            "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "achieved_GBps": step_bytes / (ms_per_step * 1e-3) / 1e9,
                              "frac_of_peak": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                              "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
                              "fwd_GBps_algorithmic": 2 * E * 4 / (fwd_ms * 1e-3) / 1e9},
            "gpu_launches": launches, "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if parity is not None:
            line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            res = cpu_reference_timing(n, c, h, w, steps=3, warmup=1, threads=os.cpu_count() or 1, budget_s=20.0)
            line["cpu_baseline"] = {
                "value": res["samples_per_s"], "unit": UNIT, "cores": res["threads"], "kind": res["kind"],
                "sample": f"{res['iters']} timed steps (mean) of the full batch {n}x{c}x{h}x{w} fp32 after 1 warm-up: " + res["what"],
                "ms_per_step": res["seconds_per_step"] * 1e3}
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that hold captured NCCL kernels must be gone before the communicator is torn down (a live graph made
        # destroy_process_group() hang on the 2-GPU box); release them, give the teardown 10 s, then leave without running
        # any more destructors.
        import gc
        import threading
        barrier()
        if gstep is not None:
            gstep.close()
        gstep = None
        gc.collect()
        torch.cuda.synchronize()
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(10.0)
        sys.stdout.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
