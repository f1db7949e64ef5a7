"""Adam over many steps: `maxstyle_step` (the arithmetic of the fused backward epilogue; bias corrections as
-expm1f(t * log1pf(-(1 - beta))) in fp32) against torch.optim.Adam (bias corrections in double on the host) fed the SAME
gradient sequence for 400 steps -- far beyond the n_iter = 5 of the reference's loop (model:537-562)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_400_adam_steps_track_torch_adam():
    from maxstyle_b200 import _lib as L, functional as F
    dev = torch.device("cuda:0")
    n, c = 20, 64
    g0 = torch.Generator(device=dev).manual_seed(3)
    params = [torch.randn(n, c, device=dev, generator=g0), torch.randn(n, c, device=dev, generator=g0), torch.rand(n, device=dev, generator=g0)]
    ours = [p.clone() for p in params]
    ref = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam(ref, lr=0.1)
    state = F.FusedStepState(F.StepConfig(), ours[0], ours[1], ours[2], True, True)
    lib = L.get_lib()
    stream = torch.cuda.current_stream().cuda_stream
    worst = 0.0
    for t in range(1, 401):
        # gradients of very different scales, some tiny (Adam's 1 / (sqrt(v) + eps) is most sensitive there)
        grads = [torch.randn(n, c, device=dev, generator=g0) * (10.0 ** ((t % 7) - 4)), torch.randn(n, c, device=dev, generator=g0),
                 torch.randn(n, device=dev, generator=g0) * 1e-3]
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        opt.step()
        st = state.struct(ours[0], ours[1], ours[2])
        rc = lib.maxstyle_step(grads[0].data_ptr(), grads[1].data_ptr(), grads[2].data_ptr(), C.byref(st), n, c, stream)
        assert rc == 0
        if t in (1, 5, 50, 400):
            torch.cuda.synchronize()
            for a, b in zip(ours, ref):
                d = float((a - b.detach()).abs().max())
                worst = max(worst, d)
                # parameters have moved by up to t * lr; fp32 rounding of ~t updates of size lr stays below 1e-5
                assert d <= 2e-5, f"step {t}: max |diff| {d:.3e}"
    assert int(state.step_dev.item()) == 400
    print(f"400 steps: max |fused - torch.optim.Adam| = {worst:.3e}")
