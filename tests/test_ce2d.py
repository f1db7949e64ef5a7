"""Pixel-wise cross entropy: oracle vs goldens generated from the reference's cross_entropy_2D on CPU;
the CUDA kernels vs the same goldens and the float64 oracle on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ce2d_oracle as CO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "CE2D_MANIFEST.json")) as _f:
    CASES = json.load(_f)["cases"]


with open(os.path.join(GOLDEN, "CE2D_SOFT_MANIFEST.json")) as _f:
    SOFT_CASES = json.load(_f)["cases"]


def _case(golden, c, name="ce2d"):
    g = golden[name]
    pre = f"c{c['idx']}_"
    return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("c", CASES, ids=lambda c: str(c["idx"]))
def test_oracle_matches_reference_cross_entropy(golden, c):
    rec = _case(golden, c)
    loss = CO.cross_entropy_2d(rec["logits"], rec["target"], c["weights"], c["size_average"], rec.get("mask"))
    grad = CO.cross_entropy_2d_grad(rec["logits"], rec["target"], c["weights"], c["size_average"], rec.get("mask"), dloss=c["dloss"])
    assert abs(loss - rec["loss"]) <= 2e-6 * max(abs(rec["loss"]), 1e-6) + 1e-7
    assert np.abs(grad - rec["dlogits"]).max() <= 2e-6 * max(np.abs(rec["dlogits"]).max(), 1e-12) + 1e-9


@pytest.mark.parametrize("c", SOFT_CASES, ids=lambda c: f"soft{c['idx']}")
def test_oracle_matches_reference_soft_target_branch(golden, c):
    """custom_loss.py:1079-1102 (4-d target: logits, or probabilities with is_gt) -- loss, d/d logits and d/d target."""
    rec = _case(golden, c, "ce2d_soft")
    kw = dict(weight=c["weights"], size_average=c["size_average"], mask=rec.get("mask"), is_gt=c["is_gt"])
    loss = CO.cross_entropy_2d_soft(rec["logits"], rec["target"], **kw)
    dl, dt = CO.cross_entropy_2d_soft_grad(rec["logits"], rec["target"], dloss=c["dloss"], **kw)
    assert abs(loss - rec["loss"]) <= 5e-6 * max(abs(rec["loss"]), 1e-6) + 1e-7
    assert np.abs(dl - rec["dlogits"]).max() <= 5e-6 * max(np.abs(rec["dlogits"]).max(), 1e-12) + 1e-9
    assert np.abs(dt - rec["dtarget"]).max() <= 5e-6 * max(np.abs(rec["dtarget"]).max(), 1e-12) + 1e-9


def test_cross_entropy_error_behaviour_on_cpu():
    from maxstyle_b200.losses import cross_entropy_2D
    with pytest.raises(RuntimeError, match="no CPU path"):
        cross_entropy_2D(torch.randn(2, 3, 4, 4), torch.randn(2, 3, 4, 4))          # soft-target branch: CUDA only, like the rest
    with pytest.raises(RuntimeError, match="does not match"):
        cross_entropy_2D(torch.randn(2, 3, 4, 4), torch.randn(2, 5, 4, 4))
    with pytest.raises(NotImplementedError):
        cross_entropy_2D(torch.randn(2, 3, 4, 4), torch.zeros(2, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        cross_entropy_2D(torch.randn(2, 3, 4, 4), torch.zeros(2, 4, 4, dtype=torch.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=lambda c: str(c["idx"]))
def test_cuda_cross_entropy_matches_reference(golden, c):
    from maxstyle_b200.losses import cross_entropy_2D
    rec = _case(golden, c)
    x = torch.from_numpy(rec["logits"]).cuda().requires_grad_(True)
    mask = None if "mask" not in rec else torch.from_numpy(rec["mask"]).cuda()
    losses = []
    for _ in range(2):                                              # twice: deterministic, workspace handling
        x.grad = None
        loss = cross_entropy_2D(x, torch.from_numpy(rec["target"]).cuda(), weight=c["weights"], size_average=c["size_average"], mask=mask)
        assert loss.dim() == 0
        (loss * c["dloss"]).backward()
        losses.append(float(loss))
    assert losses[0] == losses[1]
    # 1e-5 relative on the loss, 1e-4 on the gradient (BASELINE.json's tolerances for the path), against reference and oracle
    assert abs(losses[0] - float(rec["loss"])) <= 1e-5 * max(abs(float(rec["loss"])), 1e-6) + 1e-7
    assert _rel(x.grad.cpu().numpy(), rec["dlogits"]) < 1e-4 or np.abs(rec["dlogits"]).max() == 0
    l64 = CO.cross_entropy_2d(rec["logits"], rec["target"], c["weights"], c["size_average"], rec.get("mask"))
    g64 = CO.cross_entropy_2d_grad(rec["logits"], rec["target"], c["weights"], c["size_average"], rec.get("mask"), dloss=c["dloss"])
    assert abs(losses[0] - l64) <= 1e-5 * max(abs(l64), 1e-6) + 1e-7
    assert np.abs(x.grad.cpu().numpy() - g64).max() <= 1e-4 * max(np.abs(g64).max(), 1e-12) + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("c", SOFT_CASES, ids=lambda c: f"soft{c['idx']}")
def test_cuda_cross_entropy_soft_target_matches_reference(golden, c):
    from maxstyle_b200.losses import cross_entropy_2D
    rec = _case(golden, c, "ce2d_soft")
    mask = None if "mask" not in rec else torch.from_numpy(rec["mask"]).cuda()
    for target_grad in (True, False):
        x = torch.from_numpy(rec["logits"]).cuda().requires_grad_(True)
        t = torch.from_numpy(rec["target"]).cuda().requires_grad_(target_grad)
        loss = cross_entropy_2D(x, t, weight=c["weights"], size_average=c["size_average"], mask=mask, is_gt=c["is_gt"])
        (loss * c["dloss"]).backward()
        assert abs(float(loss) - float(rec["loss"])) <= 1e-5 * max(abs(float(rec["loss"])), 1e-6) + 1e-7
        assert _rel(x.grad.cpu().numpy(), rec["dlogits"]) < 1e-4
        if target_grad:
            assert _rel(t.grad.cpu().numpy(), rec["dtarget"]) < 1e-4
        else:
            assert t.grad is None


@pytest.mark.gpu
def test_cuda_cross_entropy_more_classes_than_registers_and_no_grad():
    """C > 8 takes the two-kernel path; a loss nobody differentiates computes no gradient."""
    from maxstyle_b200.losses import cross_entropy_2D
    import torch.nn.functional as TF
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(3, 11, 20, 24, device="cuda", generator=g).requires_grad_(True)
    t = torch.randint(0, 11, (3, 20, 24), device="cuda", generator=g)
    loss = cross_entropy_2D(x, t)
    loss.backward()
    ref = TF.cross_entropy(x.detach().requires_grad_(True), t)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    with torch.no_grad():
        plain = cross_entropy_2D(torch.randn(2, 4, 8, 8, device="cuda"), torch.randint(0, 4, (2, 8, 8), device="cuda"))
    assert plain.dim() == 0 and not plain.requires_grad


@pytest.mark.gpu
def test_cuda_cross_entropy_full_size_against_torch_and_bf16():
    """Config-2 size (20 x 4 x 224 x 224): against torch's own F.cross_entropy on the GPU; bf16 logits within bf16 resolution."""
    from maxstyle_b200.losses import cross_entropy_2D
    import torch.nn.functional as TF
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(20, 4, 224, 224, device="cuda", generator=g) * 2).requires_grad_(True)
    t = torch.randint(0, 4, (20, 224, 224), device="cuda", generator=g)
    ours = cross_entropy_2D(x, t)
    ours.backward()
    g_ours = x.grad.clone()
    x.grad = None
    ref = TF.cross_entropy(x, t)
    ref.backward()
    assert abs(float(ours) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float((g_ours - x.grad).abs().max()) <= 1e-4 * float(x.grad.abs().max())
    xb = x.detach().bfloat16().requires_grad_(True)
    lb = cross_entropy_2D(xb, t)
    lb.backward()
    ref_b = TF.cross_entropy(xb.detach().float(), t)
    assert abs(float(lb) - float(ref_b)) <= 1e-4 * abs(float(ref_b))
    assert xb.grad.dtype == torch.bfloat16
