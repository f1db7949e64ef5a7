#!/usr/bin/env python
"""Key numbers of an .ncu-rep (read here, without a GPU): duration, DRAM bytes, throughputs, stall reasons, and the source
lines that collect the most stall samples.   python tools/ncu_summary.py <file.ncu-rep> [--lines N]"""
import csv
import subprocess
import sys


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 12
    hdr, units, rows = raw(path)
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
    for r in rows:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("==", name[:100])
        for w in want:
            if w in hdr:
                print(f"   {w:72s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        print("   stalls:", ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        h = rows[0]
        col = next((i for i, x in enumerate(h) if x.strip().startswith("# Samples") or "Sampling Data (All)" in x), None)
        srccol = next((i for i, x in enumerate(h) if x.strip() == "Source"), None)
        if col is not None and srccol is not None:
            data = []
            for r in rows[1:]:
                try:
                    data.append((float(r[col].replace(",", "")), r[srccol].strip()[:110], r[0]))
                except (ValueError, IndexError):
                    pass
            tot = sum(d[0] for d in data) or 1.0
            print(f"   hottest instructions / lines ({h[col].strip()}):")
            for v, txt, addr in sorted(data, key=lambda d: -d[0])[:nlines]:
                print(f"     {100 * v / tot:5.1f}%  {txt}")


if __name__ == "__main__":
    main()
