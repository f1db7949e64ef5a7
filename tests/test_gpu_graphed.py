"""GraphedLayerStep (two CUDA graphs over static buffers) must reproduce the eager module path step by step:
same y, dX, and the same parameters after every fused Adam step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", [torch.contiguous_format, torch.channels_last])
@pytest.mark.parametrize("shape", [(6, 8, 96, 96), (20, 4, 160, 160), (5, 3, 33, 31)])
def test_graphed_step_matches_eager_module(shape, fmt):
    from maxstyle_b200 import MaxStyle, FusedStyleOptimizer, GraphedLayerStep
    n, c, h, w = shape
    g = torch.Generator().manual_seed(n + c)
    xs = [(torch.randn(shape, generator=g) * 1.4 + 0.3).cuda().contiguous(memory_format=fmt) for _ in range(4)]
    dys = [torch.randn(shape, generator=g).cuda().contiguous(memory_format=fmt) for _ in range(4)]

    torch.manual_seed(3)
    eager = MaxStyle(n, c, p=1.0)
    opt = FusedStyleOptimizer([eager], lr=0.1)
    want = []
    for x, dy in zip(xs, dys):
        xr = x.clone().requires_grad_(True)
        y = eager(xr)
        y.backward(dy)
        opt.step()
        want.append((y.detach().clone(), xr.grad.clone(), eager.lmda.detach().clone(), eager.gamma_noise.detach().clone()))

    torch.manual_seed(3)
    layer = MaxStyle(n, c, p=1.0)
    FusedStyleOptimizer([layer], lr=0.1)
    sx, sdy = xs[0].clone(), dys[0].clone()
    gs = GraphedLayerStep(layer, sx, sdy)
    # construction ran the module's first forward (batch std of xs[0], cached like the reference does) and a warm-up
    # forward, but no backward: parameters are untouched
    for i, (x, dy) in enumerate(zip(xs, dys)):
        sx.copy_(x); sdy.copy_(dy)
        y, dx = gs.run()
        for name, a, b in zip(("y", "dx", "lmda", "gamma_noise"), (y, dx, layer.lmda.detach(), layer.gamma_noise.detach()), want[i]):
            assert torch.equal(a, b), f"step {i} {name}: graphed differs from eager by {float((a.float() - b.float()).abs().max()):.3e}"
    assert int(layer._fused_step.step_dev.item()) == 4
    assert gs.kernels_per_step in (2, 4)


def test_graphed_step_without_fused_step_returns_gradients():
    from maxstyle_b200 import MaxStyle, GraphedLayerStep
    n, c, h, w = 6, 8, 64, 64
    torch.manual_seed(1)
    layer = MaxStyle(n, c, p=1.0)
    x = torch.randn(n, c, h, w, device="cuda") * 1.2 + 0.1
    dy = torch.randn(n, c, h, w, device="cuda")
    gs = GraphedLayerStep(layer, x, dy)
    y, dx = gs.run()
    xr = x.clone().requires_grad_(True)
    yr = layer(xr)
    yr.backward(dy)
    assert torch.equal(y, yr.detach()) and torch.equal(dx, xr.grad)
    for got, want in zip(gs.grads, (layer.gamma_noise.grad, layer.beta_noise.grad, layer.lmda.grad)):
        assert torch.equal(got.reshape(-1), want.reshape(-1))
    inactive = MaxStyle(n, c, p=0.0)
    with pytest.raises(RuntimeError, match="inactive"):
        GraphedLayerStep(inactive, x, dy)
