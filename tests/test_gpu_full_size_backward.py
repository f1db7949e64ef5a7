"""Full-size parity of the BENCHMARKED path against the float64 oracle (the check bench.py prints as `parity`): every element
of y and dX and every entry of d_gamma / d_beta / d_lmda on BASELINE config 1 (20x64x224x224, fp32), and 1 MiB planes
(config 5's 512x512) on a batch small enough for a float64 numpy pass.  Norm: max|a-b| / max|b| over the tensor."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(20, 64, 224, 224), (6, 4, 512, 512), (32, 16, 96, 96), (32, 1, 192, 192)])
def test_benchmarked_step_matches_fp64_oracle(shape):
    import bench
    from maxstyle_b200 import MaxStyle, FusedStyleOptimizer, GraphedLayerStep
    n, c, h, w = shape
    dev = torch.device("cuda:0")
    seed = 1234
    torch.manual_seed(seed)
    layer = MaxStyle(n, c, p=1.0)
    FusedStyleOptimizer([layer], lr=0.1, mode="adam")
    gen = torch.Generator(device=dev).manual_seed(100)
    x = torch.randn(n, c, h, w, device=dev, generator=gen) * 1.5 + 0.25
    dy = torch.randn(n, c, h, w, device=dev, generator=gen)
    gstep = GraphedLayerStep(layer, x, dy)
    for _ in range(3):                                   # first forward (batch std) + steady-state replays with the fused step
        gstep.forward()
        gstep.backward()
    out = bench.parity_check(layer, gstep, 1, 0, dev, seed)
    print(shape, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in out.items() if k not in ("tolerance", "oracle")})
    assert out["perm_exact"]
    assert out["y"] <= 1e-5, out
    # standard deviations of N nearly equal fp32 numbers (per-plane sigmas of large planes agree to 3 digits): conditioning, not a
    # kernel error -- same tolerance as the gradients
    assert out["gamma_std"] <= 1e-4 and out["beta_std"] <= 1e-4, out
    for k in ("dx", "d_gamma", "d_beta", "d_lmda"):
        assert out[k] <= 1e-4, (k, out)
    assert out["d_lmda_max_abs"] > 0.0                   # the mixing-weight gradient was exercised, not 0 == 0
    assert out["ok"]
    gstep.close() if hasattr(gstep, "close") else None
