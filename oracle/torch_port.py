"""CPU timing baseline: the reference layer's op chain restated on torch CPU tensors.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/maxstyle_oracle.py for the checker).  The reference
(src/advanced/maxstyle.py:157-185) is a chain of eager ATen ops differentiated by autograd and
stepped by torch.optim.Adam (advanced_triplet_recon_segmentation_model.py:537,562).  The
reference itself is Python and cannot travel to the GPU box, so bench.py's `cpu_baseline` and
`--impl reference` legs time THIS port: the same ATen ops in the same order on the host cores
(kind = "port").  It is validated against the reference-generated goldens by
tests/test_oracle_golden.py::test_torch_port_matches_goldens.
"""
from __future__ import annotations

import time

import torch


class StylePort:
    """State + forward of one layer on CPU tensors ([N,C,1,1] / [N,1,1,1] shapes as in the reference)."""

    def __init__(self, perm, gamma_noise, beta_noise, lmda, mix_style=True, no_noise=False, eps=1e-6):
        self.perm = torch.as_tensor(perm, dtype=torch.int64)
        n, c = gamma_noise.shape[0], gamma_noise.shape[1]
        self.gamma_noise = torch.as_tensor(gamma_noise, dtype=torch.float32).reshape(n, c, 1, 1).clone().requires_grad_(True)
        self.beta_noise = torch.as_tensor(beta_noise, dtype=torch.float32).reshape(n, c, 1, 1).clone().requires_grad_(True)
        self.lmda = torch.as_tensor(lmda, dtype=torch.float32).reshape(n, 1, 1, 1).clone().requires_grad_(True)
        self.mix_style, self.no_noise, self.eps = mix_style, no_noise, eps
        self.gamma_std = self.beta_std = None

    def parameters(self):
        return [self.gamma_noise, self.beta_noise, self.lmda]

    def forward(self, x):
        # instance statistics, detached (maxstyle.py:157-160)
        mu = x.mean(dim=[2, 3], keepdim=True)
        sig = (x.var(dim=[2, 3], keepdim=True) + self.eps).sqrt()
        mu, sig = mu.detach(), sig.detach()
        normed = (x - mu) / sig                                            # :161
        if self.gamma_std is None:                                         # :165-168, cached
            self.gamma_std = torch.std(sig, dim=0, keepdim=True).detach()
            self.beta_std = torch.std(mu, dim=0, keepdim=True).detach()
        if self.mix_style:                                                 # :172-176
            lam = torch.clamp(self.lmda, 0, 1)
            sig_mix = sig * (1 - lam) + sig[self.perm] * lam
            mu_mix = mu * (1 - lam) + mu[self.perm] * lam
        else:
            sig_mix, mu_mix = sig, mu
        if self.no_noise:                                                  # :181-182
            return sig_mix * normed + mu_mix
        return (sig_mix + self.gamma_noise * self.gamma_std) * normed + (mu_mix + self.beta_noise * self.beta_std)  # :184-185


def time_cpu_baseline(n, c, h, w, budget_s=12.0, min_iters=3, warmup=1, seed=0, threads=None):
    """Times fwd + bwd (dy given) + Adam(lr=0.1).step() of the port on the host cores.
    Returns dict(seconds_per_step (min), median, iters, threads, samples_per_s)."""
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, h, w, generator=g).requires_grad_(True)
    dy = torch.randn(n, c, h, w, generator=g)
    perm = torch.roll(torch.arange(n), 1)
    port = StylePort(perm, torch.randn(n, c, generator=g), torch.randn(n, c, generator=g), torch.rand(n, generator=g))
    opt = torch.optim.Adam(port.parameters(), lr=0.1)
    times = []
    t_begin = time.perf_counter()
    it = 0
    while True:
        t0 = time.perf_counter()
        opt.zero_grad()
        x.grad = None
        y = port.forward(x)
        y.backward(dy)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        it += 1
        if len(times) >= min_iters and time.perf_counter() - t_begin > budget_s:
            break
        if len(times) >= 50:
            break
    times.sort()
    best, med = times[0], times[len(times) // 2]
    return dict(seconds_per_step=best, median_seconds_per_step=med, iters=len(times),
                threads=torch.get_num_threads(), samples_per_s=n / best)
