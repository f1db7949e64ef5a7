"""Timing of the UNMODIFIED reference layer on the host cores.  TEST / BENCH INFRASTRUCTURE ONLY.

`time_reference_layer` runs the reference's own `MaxStyle(use_gpu=False)` (src/advanced/maxstyle.py:6-189, imported from
oracle/_ref or /root/reference through oracle/ref_shims.py) through exactly what its caller does per iteration
(model:540-566): zero_grad, forward, backward, `torch.optim.Adam(lr=0.1).step()`.  bench.py's `--impl reference` arm and its
`cpu_baseline` leg call this; when the reference is not staged they fall back to the validated port (oracle/torch_port.py).
"""
from __future__ import annotations

import time

import torch

from . import ref_shims


def time_reference_layer(n, c, h, w, steps=5, warmup=1, seed=0, threads=None, budget_s=None):
    """Exactly `steps` timed steps after `warmup` untimed ones (stops early only when `budget_s` is exceeded after >= 1 timed
    step).  Returns dict(seconds_per_step (mean over the timed steps), best, iters, threads, samples_per_s, kind)."""
    ref = ref_shims.load()
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    layer = ref.MaxStyle(n, c, p=1.0, use_gpu=False)            # p=1: the layer is active, like the GPU arm's
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, h, w, generator=g).mul_(1.5).add_(0.25).requires_grad_(True)
    dy = torch.randn(n, c, h, w, generator=g)
    opt = torch.optim.Adam(layer.parameters(), lr=0.1)
    times = []
    t_begin = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        x.grad = None
        y = layer(x)
        y.backward(dy)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            if budget_s is not None and time.perf_counter() - t_begin > budget_s:
                break
    mean = sum(times) / len(times)
    return dict(seconds_per_step=mean, best_seconds_per_step=min(times), iters=len(times), threads=torch.get_num_threads(),
                samples_per_s=n / mean, kind="reference", root=ref.root)
