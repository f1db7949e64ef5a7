#!/usr/bin/env python
"""Turn a gpurun_out/<tag>/ capture (launches.csv from `ncu --metrics gpu__time_duration.sum`, prof.ncu-rep from
`ncu --set full`) into the tracked summaries under profiles/:  <name>_launches.txt, <name>_ncu.txt, traffic.json.

    python tools/summarize_profile.py gpurun_out/r01c r01
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("ms::", "")


def launches(src, out):
    lines = [l for l in open(os.path.join(src, "launches.csv")) if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif row["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        k = short(row["Kernel Name"])
        agg.setdefault(k, []).append(v)
        order.append((k, v))
    total = sum(sum(v) for v in agg.values())
    mine = {k: v for k, v in agg.items() if "_kernel" in k and "at::" not in k}
    mine_total = sum(sum(v) for v in mine.values())
    with open(out, "w") as f:
        f.write("ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised -> compare SHARES)\n")
        f.write(f"command: python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity   (source: {src}/launches.csv)\n\n")
        f.write(f"{'kernel':70s} {'n':>4s} {'mean_us':>10s} {'total_us':>10s} {'share_all':>9s} {'share_ours':>10s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            so = f"{100 * sum(v) / mine_total:9.1f}%" if k in mine else " " * 10
            f.write(f"{k[:70]:70s} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / 1e3:10.1f} {100 * sum(v) / total:8.1f}% {so}\n")
        f.write(f"\nall kernels: {total / 1e3:.1f} us; maxstyle kernels: {mine_total / 1e3:.1f} us\n")
    return {k: sum(v) / len(v) for k, v in mine.items()}


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


REPORTS = {   # capture -> the command it was taken from (tools/gpu_profile.sh)
    "prof.ncu-rep": "python bench.py --steps 3 --warmup 5 --no-e2e --no-cpu-baseline --no-parity   (-k regex:fwd_pair|bwd_nchw -s 12 -c 2)",
    "prof2.ncu-rep": "MAXSTYLE_SWEEP=18,3,4 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline   (two-pass forward forced)",
    "prof3.ncu-rep": "python tools/kernel_bench.py --fwd-only --dtype bf16 --sweeps 2,3,4 --iters 10   (bf16 config-1 shape)",
    "prof4.ncu-rep": "python tools/kernel_bench.py --layout nhwc --sweeps 2,2,4   (NHWC kernels)",
}


def ncu(src, out, traffic_path):
    traffic = {}
    with open(out, "w") as f:
        f.write("ncu --set full --clock-control none --import-source on  (one launch per kernel; cold caches)\n")
        for rep_name, cmd in REPORTS.items():
            rep = os.path.join(src, rep_name)
            if os.path.exists(rep):
                ncu_one(rep, cmd, f, traffic)
    json.dump(traffic, open(traffic_path, "w"), indent=1)


def ncu_one(rep, cmd, f, traffic):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
    if True:
        f.write(f"\ncommand: {cmd}   (source: {rep})\n")
        for r in rows[2:]:
            name = short(r[idx["Kernel Name"]])
            f.write(f"\n== {name}\n")
            for w in WANT:
                if w in idx:
                    f.write(f"   {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}\n")
            tot = sum(float(r[idx[h]].replace(",", "") or 0) for h in stall) or 1.0
            top = sorted(((float(r[idx[h]].replace(",", "") or 0), h) for h in stall), reverse=True)[:5]
            f.write("   top stall reasons (pc sampling): " + ", ".join(f"{h[33:]} {100 * v / tot:.0f}%" for v, h in top) + "\n")
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            key = name.split("<")[0]
            if key in traffic:
                key = name
            traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": name,
                            "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]].replace(",", ""))}


def main():
    src, name = sys.argv[1], sys.argv[2]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(src, os.path.join(ROOT, "profiles", f"{name}_launches.txt"))
    ncu(src, os.path.join(ROOT, "profiles", f"{name}_ncu.txt"), os.path.join(ROOT, "profiles", "traffic.json"))
    for extra in ("bench.json", "bench_ref.json", "pytest_gpu.txt", "smoke.txt", "fwd_paths.txt", "fwd_paths.jsonl", "configs.jsonl", "sweep.jsonl",
                  "loop_config2.txt", "loop_config2_ref.txt", "ring_knobs.txt", "bench_config3.json", "bench_config5.json", "ce2d.txt",
                  "pm_series_bench.txt"):
        p = os.path.join(src, extra)
        if os.path.exists(p):
            with open(p) as fi, open(os.path.join(ROOT, "profiles", f"{name}_{extra}"), "w") as fo:
                fo.write(fi.read())


if __name__ == "__main__":
    main()
