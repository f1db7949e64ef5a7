#!/usr/bin/env python
"""Forward-only timings of maxstyle_fwd per path (CUDA events, steady state = cached batch std) with the cluster
kernel's geometry printed beside each line.  Development tool.

    python tools/cluster_bench.py [--shapes "20,64,224,224,f32;..."] [--iters 30]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="20,64,224,224,f32;20,64,224,224,bf16;64,64,112,112,f32;32,16,192,192,f32;32,16,96,96,f32;"
                                        "32,1,192,192,f32;20,16,224,224,f32;64,32,512,512,f32")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--variants", default="")
    args = ap.parse_args()
    from maxstyle_b200 import functional as F, _lib as L, MaxStyle
    lib = L.get_lib()
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    FC = L.SWEEP_FORCE_CLUSTER
    variants = [("default", 0), ("two_pass", L.SWEEP_NO_FUSED), ("window", L.SWEEP_NO_RESIDENT | L.SWEEP_NO_RING | L.SWEEP_FORCE_WINDOW),
                ("resident", L.SWEEP_FORCE_RESIDENT), ("cluster_auto", FC)]
    for pc in (0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 12, 16, 24, 32):
        variants.append((f"pair_p{pc}", L.SWEEP_FORCE_PAIR | (pc << L.SWEEP_CLUSTER_PIECES_SHIFT)))
    variants.append(("pair_p0_normal", L.SWEEP_FORCE_PAIR | L.SWEEP_X_STREAM))
    variants.append(("pair_p4_normal", L.SWEEP_FORCE_PAIR | L.SWEEP_X_STREAM | (4 << L.SWEEP_CLUSTER_PIECES_SHIFT)))

    def geo_sweep(cs, stages, pieces):
        return FC | (cs << L.SWEEP_CLUSTER_SIZE_SHIFT) | (stages << L.SWEEP_CLUSTER_STAGES_SHIFT) | (pieces << L.SWEEP_CLUSTER_PIECES_SHIFT)
    for cs, st, pc in ((1, 0, 1), (2, 0, 1), (1, 0, 2), (2, 0, 2)):
        variants.append((f"cluster_cs{cs}_p{pc}" + (f"_s{st}" if st else ""), geo_sweep(cs, st, pc)))
    if args.variants:
        keep = set(args.variants.split(","))
        variants = [v for v in variants if v[0] in keep]
    for spec in args.shapes.split(";"):
        n, c, h, w, dts = spec.split(",")
        n, c, h, w = int(n), int(c), int(h), int(w)
        dt = torch.float32 if dts == "f32" else torch.bfloat16
        es = 4 if dts == "f32" else 2
        torch.manual_seed(0)
        x = (torch.randn(n, c, h, w, device=dev) * 1.5 + 0.25).to(dt)
        y = torch.empty_like(x)
        layer = MaxStyle(n, c, p=1.0)
        code = F.dtype_code(x)
        ws = F.new_workspace(n, c, h, w, code, dev)
        perm = layer.perm.to(dev)
        gs = torch.empty(c, device=dev); bs = torch.empty(c, device=dev)
        tabs = torch.empty(4, n, c, device=dev)
        lm, gn, bn = layer.lmda.detach(), layer.gamma_noise.detach(), layer.beta_noise.detach()
        E = x.numel()
        ref = None
        for name, sweep in variants:
            if lib.maxstyle_fwd_kernels(n, c, h, w, code, L.NCHW, sweep) != 1 and name not in ("default", "two_pass"):
                continue
            geo = (C.c_int * 12)()
            geom = None
            if name.startswith("cluster") or name.startswith("pair") or name == "default":
                if lib.maxstyle_fwd_geometry(n, c, h, w, code, sweep, geo) == 0:
                    geom = dict(cs=geo[0], stages=geo[1], clusters=geo[2], part=geo[3], chunk=geo[4], chunks=geo[5], smem=geo[6], pieces=geo[8], order=geo[9])
            F.SWEEP_STATS = sweep
            try:
                F.forward_raw(x, perm, lm, gn, bn, gs, bs, L.FLAG_MIX_STYLE | L.FLAG_COMPUTE_BATCH_STD, 1e-6, ws, out=y, tables=tabs)
                for _ in range(5):
                    F.forward_raw(x, perm, lm, gn, bn, gs, bs, L.FLAG_MIX_STYLE, 1e-6, ws, out=y, tables=tabs)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    F.forward_raw(x, perm, lm, gn, bn, gs, bs, L.FLAG_MIX_STYLE, 1e-6, ws, out=y, tables=tabs)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / args.iters * 1e3
                # first-forward variant too
                e0.record()
                for _ in range(10):
                    F.forward_raw(x, perm, lm, gn, bn, gs, bs, L.FLAG_MIX_STYLE | L.FLAG_COMPUTE_BATCH_STD, 1e-6, ws, out=y, tables=tabs)
                e1.record()
                torch.cuda.synchronize()
                us_first = e0.elapsed_time(e1) / 10 * 1e3
                F.workspace_status(ws, n, c, h, w, code)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"shape": [n, c, h, w], "dtype": dts, "path": name, "error": str(e)[:200]}), flush=True)
                raise
            chk = float(y.float().abs().mean())
            if ref is None:
                ref = chk
            print(json.dumps({"shape": [n, c, h, w], "dtype": dts, "path": name, "fwd_us": round(us, 1), "first_fwd_us": round(us_first, 1),
                              "GBps_2E": round(2 * E * es / us / 1e3), "frac": round(2 * E * es / us / 1e3 / peak, 3),
                              "geometry": geom, "check": round(chk / ref, 6)}), flush=True)


if __name__ == "__main__":
    main()
