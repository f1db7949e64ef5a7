#!/usr/bin/env python
"""Probe: does torch's symmetric memory (peer pointers over NVLink) work on this box?  Run under torchrun."""
import os, json
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"rank": rank}
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(4096, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, group=dist.group.WORLD)
    out["buffer_ptrs"] = [hex(p) for p in hdl.buffer_ptrs]
    out["signal_pad_ptrs"] = [hex(p) for p in hdl.signal_pad_ptrs]
    out["signal_pad_size"] = getattr(hdl, "signal_pad_size", None)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (4096,), torch.float32)
    out["peer_value"] = float(peer[0].item())
    hdl.barrier()
    out["ok"] = True
except Exception as e:  # noqa: BLE001
    out["ok"] = False
    out["error"] = repr(e)[:400]
print(json.dumps(out), flush=True)
dist.barrier()
os._exit(0)
