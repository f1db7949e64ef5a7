#!/usr/bin/env python
"""BASELINE.json config 4: layer sweep for HBM-roofline characterisation.

N in {64..512}, C in {16..256}, HxW in {28^2..224^2}, fp32 and bf16, NCHW and NHWC.
Per point: forward, backward (+ fused Adam step) and whole-step time through the C ABI (CUDA events on the
launching stream, median), achieved algorithmic GB/s (5*E*s per step) and its fraction of the measured HBM peak.
Points whose four live tensors fit the 126 MB L2 get an L2 flush (a 512 MB fill) before every timed
iteration and are marked `fits_l2`; they are not HBM-roofline evidence without it.

    python tools/sweep.py [--quick] [--max-gb 8] [--out profiles/r01_sweep.jsonl]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6553.0, "fallback (round-1 measured value)"


def point(n, c, h, w, dtype, layout, iters, flush_buf, peak):
    layout = layout.upper()
    from maxstyle_b200 import functional as F, _lib as L, MaxStyle, FusedStyleOptimizer
    dev = torch.device("cuda:0")
    dt = torch.float32 if dtype == "f32" else torch.bfloat16
    es = 4 if dtype == "f32" else 2
    fmt = torch.contiguous_format if layout == "NCHW" else torch.channels_last
    g = torch.Generator(device=dev).manual_seed(hash((n, c, h, w)) & 0xffffffff)
    # per-plane scale/shift so the style tables are non-degenerate
    x = torch.randn(n, c, h, w, device=dev, generator=g)
    x.mul_(torch.rand(n, c, 1, 1, device=dev, generator=g) * 1.5 + 0.5).add_(torch.rand(n, c, 1, 1, device=dev, generator=g) * 2 - 1)
    x = x.to(dt).contiguous(memory_format=fmt)
    dy = torch.randn(n, c, h, w, device=dev, generator=g).to(dt).contiguous(memory_format=fmt)
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    E = x.numel()
    layer = MaxStyle(n, c, p=1.0)
    opt = FusedStyleOptimizer([layer], lr=0.1)
    lay = F.layout_of(x)
    ws = F.new_workspace(n, c, h, w, F.dtype_code(x), dev, lay)
    perm = layer.perm.to(dev)
    gs = torch.empty(c, device=dev); bs = torch.empty(c, device=dev)
    tabs = torch.empty(4, n, c, device=dev)
    flags = L.FLAG_MIX_STYLE
    step = layer._fused_step.struct(layer.gamma_noise, layer.beta_noise, layer.lmda)
    fits = 4 * E * es <= 126e6

    def fwd(first=False):
        F.forward_raw(x, perm, layer.lmda, layer.gamma_noise, layer.beta_noise, gs, bs,
                      flags | (L.FLAG_COMPUTE_BATCH_STD if first else 0), 1e-6, ws, out=y, tables=tabs)

    def bwd():
        F.backward_raw(dy, x, tabs[0], tabs[1], 0, tabs[2], perm, layer.lmda, gs, bs, flags, ws, dx_out=dx,
                       need_noise_grad=False, need_mix_grad=False, step=step)

    fwd(True); bwd(); fwd(); bwd()
    torch.cuda.synchronize()
    # Forward and backward are captured as two CUDA graphs and replayed: launched eagerly from Python a call costs 15-25 us of
    # host time, more than the small shapes' kernels take (the C-ABI calls are capturable: no host reads, caller's stream).
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fwd(); bwd()
        side.synchronize()
        with torch.cuda.graph(gf, stream=side):
            fwd()
        with torch.cuda.graph(gb, stream=side):
            bwd()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        gf.replay(); gb.replay()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
    for i in range(iters):
        if fits:
            flush_buf.fill_(i)
        ev[i][0].record(); gf.replay(); ev[i][1].record(); gb.replay(); ev[i][2].record()
    torch.cuda.synchronize()
    med = lambda a, b: sorted(ev[i][a].elapsed_time(ev[i][b]) for i in range(iters))[iters // 2] * 1e3
    f_us, b_us, s_us = med(0, 1), med(1, 2), med(0, 2)
    gbps = 5 * E * es / s_us / 1e3
    return {"N": n, "C": c, "H": h, "W": w, "dtype": dtype, "layout": layout, "MB_per_tensor": round(E * es / 1e6, 1),
            "fits_l2": fits, "fwd_kernels": F.fwd_kernel_count(n, c, h, w, F.dtype_code(x), lay),
            "fwd_us": round(f_us, 1), "bwd_us": round(b_us, 1), "step_us": round(s_us, 1),
            "samples_per_s": round(n / s_us * 1e6), "GBps_5E": round(gbps), "frac_hbm_peak": round(gbps / peak, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="corner + centre points only")
    ap.add_argument("--max-gb", type=float, default=8.0, help="skip points whose tensors exceed this many GB each")
    ap.add_argument("--iters", type=int, default=9)
    ap.add_argument("--out", default="")
    ap.add_argument("--points", default="", help='explicit points instead of the grid: "N,C,H,W,dtype,layout;..." '
                    '(e.g. BASELINE config 5 per GPU: 256,32,512,512,f32,NCHW)')
    args = ap.parse_args()
    peak, src = hbm_peak()
    Ns, Cs, Ss = [64, 128, 256, 512], [16, 32, 64, 128, 256], [28, 56, 112, 224]
    if args.quick:
        Ns, Cs, Ss = [64, 512], [16, 64, 256], [28, 112, 224]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    out = open(args.out, "w") if args.out else None
    head = {"sweep": "BASELINE config 4", "hbm_peak_GBps": peak, "peak_source": src, "iters": args.iters, "max_gb": args.max_gb,
            "note": "step = maxstyle_fwd + maxstyle_bwd with fused Adam through the C ABI, each captured in a CUDA graph and replayed; 5*E*s algorithmic bytes"}
    print(json.dumps(head)); out and out.write(json.dumps(head) + "\n")
    skipped = 0
    if args.points:
        for spec in args.points.split(";"):
            n, c, h, w, dtype, layout = spec.split(",")
            r = point(int(n), int(c), int(h), int(w), dtype, layout, args.iters, flush, peak)
            line = json.dumps(r)
            print(line, flush=True)
            if out:
                out.write(line + "\n"); out.flush()
            torch.cuda.empty_cache()
        return
    for dtype in ("f32", "bf16"):
        for layout in ("NCHW", "NHWC"):
            for n in Ns:
                for c in Cs:
                    for s in Ss:
                        es = 4 if dtype == "f32" else 2
                        if n * c * s * s * es > args.max_gb * 1e9:
                            skipped += 1
                            continue
                        r = point(n, c, s, s, dtype, layout, args.iters, flush, peak)
                        line = json.dumps(r)
                        print(line, flush=True)
                        if out:
                            out.write(line + "\n"); out.flush()
                        torch.cuda.empty_cache()
    tail = {"skipped_over_max_gb": skipped}
    print(json.dumps(tail)); out and out.write(json.dumps(tail) + "\n")


if __name__ == "__main__":
    main()
