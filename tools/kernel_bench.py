#!/usr/bin/env python
"""Per-kernel timings through the C ABI (CUDA events on the launching stream) + host overhead of the
eager module path.  Development tool; bench.py is the contract benchmark.

    python tools/kernel_bench.py [--shape N,C,H,W] [--dtype f32|bf16] [--iters 30] [--sweeps "2,3,4;0,0,0"]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="20,64,224,224")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--sweeps", default="2,3,4;0,0,0;2,3,0;0,1,0;2,2,4")
    ap.add_argument("--host", action="store_true", help="also measure host overhead of the module path")
    ap.add_argument("--fwd-only", action="store_true", help="only time maxstyle_fwd")
    ap.add_argument("--layout", default="nchw", choices=["nchw", "nhwc"])
    args = ap.parse_args()
    from maxstyle_b200 import functional as F, _lib as L, MaxStyle, FusedStyleOptimizer
    n, c, h, w = (int(v) for v in args.shape.split(","))
    dt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    es = 4 if args.dtype == "f32" else 2
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = (torch.randn(n, c, h, w, device=dev) * 1.5 + 0.25).to(dt)
    dy = torch.randn(n, c, h, w, device=dev).to(dt)
    if args.layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last); dy = dy.contiguous(memory_format=torch.channels_last)
    y = torch.empty_like(x); dx = torch.empty_like(x)
    E = x.numel()
    layer = MaxStyle(n, c, p=1.0)
    ws = F.new_workspace(n, c, h, w, F.dtype_code(x), dev, F.layout_of(x))
    perm = layer.perm.to(dev)
    gs = torch.empty(c, device=dev); bs = torch.empty(c, device=dev)
    tabs = torch.empty(4, n, c, device=dev)
    mu, sig, scale, shift = tabs[0], tabs[1], tabs[2], tabs[3]
    lm, gn, bn = layer.lmda.detach(), layer.gamma_noise.detach(), layer.beta_noise.detach()
    flags = L.FLAG_MIX_STYLE | L.FLAG_COMPUTE_BATCH_STD
    peak = 6553.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def run(sw, iters):
        s_st, s_ap, s_bw = sw
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(iters)]
        for i in range(iters):
            ev[i][0].record()
            F.instance_stats(x, 1e-6, ws, mu, sig, 0, sweep=s_st)
            ev[i][1].record()
            F.style_tables(mu, sig, 0, n, perm, lm, gn, bn, gs, bs, flags, scale, shift)
            ev[i][2].record()
            F.style_apply(x, mu, 0, scale, shift, out=y, sweep=s_ap)
            ev[i][3].record()
            F.SWEEP_BWD = s_bw
            F.backward_raw(dy, x, mu, sig, 0, scale, perm, lm, gs, bs, flags & 3, ws, dx_out=dx,
                           grads_out=(tabs[2].new_empty(n, c), tabs[2].new_empty(n, c), tabs[2].new_empty(n)))
            ev[i][4].record()
        torch.cuda.synchronize()
        def med(a, b):
            v = sorted(ev[i][a].elapsed_time(ev[i][b]) for i in range(iters // 3, iters))
            return v[len(v) // 2]
        return med(0, 1), med(1, 2), med(2, 3), med(3, 4), med(0, 4)

    if args.fwd_only:
        F.SWEEP_STATS, F.SWEEP_APPLY, F.SWEEP_BWD = (int(v) for v in args.sweeps.split(";")[0].split(","))
        for _ in range(5):
            F.forward_raw(x, perm, lm, gn, bn, gs, bs, flags, 1e-6, ws, out=y, tables=tabs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            F.forward_raw(x, perm, lm, gn, bn, gs, bs, flags & 3, 1e-6, ws, out=y, tables=tabs)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / args.iters * 1e3
        print(json.dumps({"shape": [n, c, h, w], "dtype": args.dtype, "fwd_us": round(us, 1), "fwd_GBps_2E": round(2 * E * es / us / 1e3),
                          "kernels": F.fwd_kernel_count(n, c, h, w, F.dtype_code(x)), "piece_kb": os.environ.get("MAXSTYLE_FUSED_PIECE_KB"),
                          "window_mb": os.environ.get("MAXSTYLE_FUSED_WINDOW_MB"), "chunk": os.environ.get("MAXSTYLE_FUSED_CHUNK")}))
        return
    for sw in args.sweeps.split(";"):
        sw = tuple(int(v) for v in sw.split(","))
        run(sw, 5)
        st, tb, ap, bw, tot = run(sw, args.iters)
        gb = lambda k, ms: k * E * es / (ms * 1e-3) / 1e9
        print(json.dumps({"shape": [n, c, h, w], "dtype": args.dtype, "layout": args.layout, "sweeps": sw,
                          "stats_us": round(st * 1e3, 1), "tables_us": round(tb * 1e3, 1), "apply_us": round(ap * 1e3, 1),
                          "bwd_us": round(bw * 1e3, 1), "step_us": round(tot * 1e3, 1),
                          "stats_GBps": round(gb(1, st)), "apply_GBps": round(gb(2, ap)), "bwd_GBps": round(gb(3, bw)),
                          "step_frac_5E": round(gb(5, tot) / peak, 3)}))

    # ---- where does the eager module path lose time against the raw sequence? ----------------------
    def timed(fn, iters=40, warm=8):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    F.SWEEP_STATS, F.SWEEP_APPLY, F.SWEEP_BWD = (int(v) for v in args.sweeps.split(";")[0].split(","))
    big = MaxStyle(n, c, p=1.0)
    xr = x.detach().clone().requires_grad_(True)
    res = {}

    def module_step():
        xr.grad = None
        big(xr).backward(dy)
    res["module_torch_grads_us"] = round(timed(module_step), 1)
    opt = FusedStyleOptimizer([big], lr=0.1)
    res["module_fused_adam_us"] = round(timed(module_step), 1)
    st_struct = big._fused_step.struct(big.gamma_noise, big.beta_noise, big.lmda)
    gstd, bstd = big.gamma_std.view(-1), big.beta_std.view(-1)

    def raw_step(step=None, fresh=False):
        yy = torch.empty_like(x) if fresh else y
        dd = torch.empty_like(x) if fresh else dx
        F.forward_raw(x, perm, big.lmda, big.gamma_noise, big.beta_noise, gstd, bstd, flags & 3, 1e-6, ws, out=yy, tables=tabs)
        F.backward_raw(dy, x, mu, sig, 0, scale, perm, big.lmda, gstd, bstd, flags & 3, ws, dx_out=dd,
                       need_noise_grad=step is None, need_mix_grad=step is None, step=step)
    res["fwd_only_us"] = round(timed(lambda: F.forward_raw(x, perm, big.lmda, big.gamma_noise, big.beta_noise, gstd, bstd, flags & 3,
                                                           1e-6, ws, out=y, tables=tabs)), 1)
    res["fwd_kernels"] = F.fwd_kernel_count(n, c, h, w, F.dtype_code(x))
    res["raw_nograd_step_us"] = round(timed(lambda: raw_step(None)), 1)
    res["raw_fused_adam_us"] = round(timed(lambda: raw_step(st_struct)), 1)
    res["raw_fused_adam_fresh_alloc_us"] = round(timed(lambda: raw_step(st_struct, True)), 1)
    # the same raw sequence replayed from a CUDA graph
    gph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        raw_step(st_struct)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(gph):
        raw_step(st_struct)
    res["graph_fused_adam_us"] = round(timed(gph.replay), 1)
    res["roofline_5E_us"] = round(5 * E * es / (peak * 1e9) * 1e6, 1)
    print(json.dumps(res))

    if args.host:
        # host overhead: tiny tensors, so the GPU is never the bottleneck; wall clock per eager step
        nn_, cc_ = 8, 16
        small = MaxStyle(nn_, cc_, p=1.0)
        opt = FusedStyleOptimizer([small], lr=0.1)
        xs = torch.randn(nn_, cc_, 32, 32, device=dev, requires_grad=True)
        dys = torch.randn(nn_, cc_, 32, 32, device=dev)
        for _ in range(20):
            small(xs).backward(dys)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 300
        for _ in range(k):
            xs.grad = None
            small(xs).backward(dys)
            opt.step()
        torch.cuda.synchronize()
        per = (time.perf_counter() - t0) / k
        t0 = time.perf_counter()
        with torch.no_grad():
            for _ in range(k):
                small(xs)
        torch.cuda.synchronize()
        per_f = (time.perf_counter() - t0) / k
        print(json.dumps({"host_us_per_eager_step": round(per * 1e6, 1), "host_us_per_forward_nograd": round(per_f * 1e6, 1)}))
        import cProfile, pstats, io
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(200):
            xs.grad = None
            small(xs).backward(dys)
            opt.step()
        torch.cuda.synchronize()
        pr.disable()
        sio = io.StringIO()
        pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(25)
        print(sio.getvalue()[:5000])


if __name__ == "__main__":
    main()
