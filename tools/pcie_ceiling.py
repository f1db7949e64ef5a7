#!/usr/bin/env python
"""What the host link gives each rank when N ranks copy at once: bare pinned-memory copies, no kernels.

    python tools/pcie_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

Per rank: H2D alone, D2H alone, and both at once (two streams), 257 MB per copy (the size of one feature map of bench.py's
workload), every rank copying at the same time (barrier before each phase).  Rank 0 prints one JSON line with the per-rank
rates, their sum, and where each process is allowed to run (CPU affinity) -- enough to tell a link limit of the box from a
problem of maxstyle_b200.HostStepPipeline, whose `e2e` number can be at most `both` each way.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 20 * 64 * 224 * 224
    h_in = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.float32, device=dev)
    d_out = torch.randn(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    nbytes = n * 4

    def phase(h2d, d2h, iters=6):
        def run(k):
            for _ in range(k):
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            s1.synchronize(); s2.synchronize()
        run(1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(iters)
        return nbytes * iters / (time.perf_counter() - t0) / 1e9

    res = [phase(True, False), phase(False, True), phase(True, True)]
    t = torch.tensor(res, dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
    else:
        allr = [t]
    aff = sorted(os.sched_getaffinity(0))
    if rank == 0:
        rows = [[round(float(v), 2) for v in r.tolist()] for r in allr]
        print(json.dumps({
            "ranks": world, "bytes_per_copy": nbytes,
            "per_rank_GBps": {"h2d_alone": [r[0] for r in rows], "d2h_alone": [r[1] for r in rows], "both_each_way": [r[2] for r in rows]},
            "sum_GBps": {"h2d_alone": round(sum(r[0] for r in rows), 1), "d2h_alone": round(sum(r[1] for r in rows), 1),
                         "both_each_way": round(sum(r[2] for r in rows), 1)},
            "cpu_affinity_rank0": f"{aff[0]}-{aff[-1]} ({len(aff)} cpus)", "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    sys.exit(main())
