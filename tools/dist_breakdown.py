#!/usr/bin/env python
"""Where the multi-GPU forward spends its time: statistics kernel, the all-gather of the (mu|sig) rows, table kernel,
apply kernel -- CUDA events on each rank, max over ranks.  Run under torchrun (one rank per GPU)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from maxstyle_b200 import GlobalBatchMaxStyle, functional as F, _lib as L
    from maxstyle_b200.distributed import StyleTableExchange
    n, c, h, w = 20, 64, 224, 224
    torch.manual_seed(0)
    layer = GlobalBatchMaxStyle(n, c, p=1.0)
    x = torch.randn(n, c, h, w, device=dev) * 1.5 + 0.25
    y = torch.empty_like(x)
    ws = F.new_workspace(n, c, h, w, L.F32, dev)
    ex = layer._exchange
    table = ex.allocate(n, c, dev)
    mu_all, sig_all = StyleTableExchange.views(table)
    gs, bs = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    perm = layer._perm_device(dev)
    iters = 40
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(iters)]
    for i in range(iters):
        ev[i][0].record()
        F.instance_stats(x, 1e-6, ws, mu_all, sig_all, layer.row_offset)
        ev[i][1].record()
        ex.gather(table, n)
        ev[i][2].record()
        scale, shift = F.style_tables(mu_all, sig_all, layer.row_offset, n, perm, layer.lmda.detach(), layer.gamma_noise.detach(),
                                      layer.beta_noise.detach(), gs, bs, L.FLAG_MIX_STYLE | L.FLAG_COMPUTE_BATCH_STD)
        ev[i][3].record()
        F.style_apply(x, mu_all, layer.row_offset, scale, shift, out=y)
        ev[i][4].record()
    torch.cuda.synchronize()
    med = lambda a, b: sorted(ev[i][a].elapsed_time(ev[i][b]) for i in range(iters // 2, iters))[iters // 4] * 1e3
    t = torch.tensor([med(0, 1), med(1, 2), med(2, 3), med(3, 4), med(0, 4)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps(dict(world=world, stats_us=round(t[0].item(), 1), allgather_us=round(t[1].item(), 1), tables_us=round(t[2].item(), 1),
                              apply_us=round(t[3].item(), 1), forward_us=round(t[4].item(), 1))))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
