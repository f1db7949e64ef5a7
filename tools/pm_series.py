#!/usr/bin/env python
"""DRAM read / write throughput over the run time of each kernel in an .ncu-rep (the PmSampling section of `ncu --set full`,
1 us samples), printed as one row per kernel and metric.  Reads the report here, without a GPU.
    python tools/pm_series.py <file.ncu-rep>"""
import glob
import sys

for d in glob.glob("/opt/nvidia/nsight-compute/*/extras/python"):
    sys.path.insert(0, d)
import ncu_report  # noqa: E402


def main():
    ctx = ncu_report.load_report(sys.argv[1])
    for ri in range(ctx.num_ranges()):
        rng = ctx.range_by_idx(ri)
        for ai in range(rng.num_actions()):
            act = rng.action_by_idx(ai)
            series = {}
            for n in act.metric_names():
                if "dram__" in n and "throughput" in n and n.split(".")[0] in ("FBSP", "FBPA"):
                    m = act.metric_by_name(n)
                    if m.num_instances() > 1:
                        series[n] = [m.as_double(i) for i in range(m.num_instances())]
            if not series:
                continue
            tot = next((v for k, v in series.items() if "dram__throughput" in k), None)
            lo = next((i for i, v in enumerate(tot) if v > 0.5), 0)
            hi = max(i for i, v in enumerate(tot) if v > 0.5) + 1
            dur = act.metric_by_name("gpu__time_duration.sum").as_double() / 1e3
            print(f"== {act.name()}  duration {dur:.1f} us, samples {lo}..{hi} (1 per us), % of ncu's DRAM peak")
            for k in sorted(series):
                kind = "read " if "read" in k else ("write" if "write" in k else "total")
                vals = series[k][lo:hi]
                print(f"   {kind} mean {sum(vals) / len(vals):5.1f} | " + " ".join(f"{v:3.0f}" for v in vals))


if __name__ == "__main__":
    main()
