"""HostStepPipeline (the host-buffer entry point bench.py's e2e line times): results in the host slots must equal
the device-resident module call step by step, with two steps in flight."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(n, c, seed):
    from maxstyle_b200 import MaxStyle, FusedStyleOptimizer
    torch.manual_seed(seed)
    layer = MaxStyle(n, c, p=1.0)
    opt = FusedStyleOptimizer([layer], lr=0.1)
    return layer, opt


def test_pipeline_matches_direct_module_calls():
    from maxstyle_b200 import HostStepPipeline
    n, c, h, w = 6, 8, 64, 64
    steps = 5
    g = torch.Generator().manual_seed(7)
    hxs = [(torch.randn(n, c, h, w, generator=g) * 1.3 + 0.2).pin_memory() for _ in range(steps)]
    hdys = [torch.randn(n, c, h, w, generator=g).pin_memory() for _ in range(steps)]

    # direct: device-resident module calls, one after the other
    layer, opt = _mk(n, c, 11)
    want = []
    for hx, hdy in zip(hxs, hdys):
        x = hx.cuda().requires_grad_(True)
        y = layer(x)
        y.backward(hdy.cuda())
        opt.step()
        params = torch.cat([layer.gamma_noise.detach().flatten(), layer.beta_noise.detach().flatten(), layer.lmda.detach().flatten()])
        want.append((y.detach().cpu(), x.grad.cpu(), params.cpu()))

    # pipelined from host buffers, same seed -> same initial parameters
    layer2, opt2 = _mk(n, c, 11)
    pipe = HostStepPipeline(layer2, (n, c, h, w), torch.float32, depth=2)
    tickets = []
    got = []
    for i in range(steps):
        tickets.append(pipe.submit(hxs[i], hdys[i]))
        if i >= 1:
            r = pipe.wait(tickets[i - 1])
            got.append((r.y.clone(), r.dx.clone(), r.params.clone()))
    r = pipe.wait(tickets[-1])
    got.append((r.y.clone(), r.dx.clone(), r.params.clone()))
    pipe.drain()
    assert pipe.h2d_bytes == 2 * n * c * h * w * 4 and pipe.d2h_bytes == pipe.h2d_bytes + (2 * n * c + n) * 4
    for i, (a, b) in enumerate(zip(got, want)):
        for nm, u, v in zip(("y", "dx", "params"), a, b):
            assert torch.equal(u, v), f"step {i} {nm}: pipeline differs from the direct call by {float((u - v).abs().max()):.3e}"


def test_pipeline_rejects_bad_buffers_and_stale_tickets():
    from maxstyle_b200 import HostStepPipeline
    layer, _ = _mk(4, 4, 3)
    pipe = HostStepPipeline(layer, (4, 4, 32, 32))
    good = torch.zeros(4, 4, 32, 32).pin_memory()
    with pytest.raises(RuntimeError, match="host tensor"):
        pipe.submit(good.cuda(), good)
    with pytest.raises(RuntimeError, match="host tensor"):
        pipe.submit(torch.zeros(4, 4, 16, 16), good)
    t = pipe.submit(good + 1.0, good)
    pipe.wait(t)
    with pytest.raises(RuntimeError, match="not in flight"):
        pipe.wait(t + 5)
