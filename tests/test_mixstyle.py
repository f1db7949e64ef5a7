"""MixStyle / DSU layer: oracle vs reference-generated goldens on CPU; the CUDA module vs the same
goldens on the GPU under the reference's seeds (the mixing weight and the permutation are drawn on the CPU generator in
the reference too, so the same seed must give the same augmentation)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mixstyle_oracle as MO
from oracle.gen_golden import make_input

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "MIXSTYLE_MANIFEST.json")) as _f:
    MANIFEST = json.load(_f)
CASES = [c for c in MANIFEST["cases"] if c["idx"] != "identity"]


def _case(golden, c):
    g = golden["mixstyle"]
    pre = f"c{c['idx']}_"
    rec = {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}
    x = make_input(c["x_seed"], tuple(c["shape"]), c["kind"])
    dy = np.random.RandomState(c["dy_seed"]).standard_normal(size=tuple(c["shape"])).astype(np.float32)
    return x, dy, rec


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"{c['idx']}-{c['mix']}")
def test_oracle_matches_reference_mixstyle(golden, c):
    x, dy, rec = _case(golden, c)
    y, cache = MO.mixstyle_forward(x, c["mix"], lmda=rec.get("lmda"), perm=rec.get("perm"), eps_mu=rec.get("eps_mu"),
                                   eps_sig=rec.get("eps_sig"), eps=1e-8, dtype=np.float64)
    # 'offset' inputs (|mu|/sig = 1e4): the reference's own fp32 normalisation carries ~1e-3 of relative noise
    tol = 5e-3 if c["kind"] == "offset" else 2e-5
    assert _rel(y, rec["y"]) < tol
    assert _rel(MO.mixstyle_backward(dy, cache, dtype=np.float64), rec["dx"]) < tol
    if c["mix"] == "crossdomain":
        assert MO.crossdomain_perm_is_valid(rec["perm"])
    if c["fixed_lmda"] is not None:
        assert np.all(rec["lmda"] == np.float32(c["fixed_lmda"])) and c["fixed_lmda"] > 1.0      # extrapolation is exercised


def test_reference_identity_branch_recorded():
    ident = [c for c in MANIFEST["cases"] if c["idx"] == "identity"][0]
    assert ident["returns_same_object"] is True


def test_cpu_input_fails_loudly():
    from maxstyle_b200 import MixStyle
    m = MixStyle(p=1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(4, 3, 8, 8))
    m0 = MixStyle(p=0.0)
    x = torch.randn(4, 3, 8, 8)
    torch.manual_seed(0)
    assert m0(x) is x                      # inactive draw: the same tensor object, like the reference
    assert repr(MixStyle(p=0.3, alpha=0.2, mix="gaussian")) == "MixStyle(p=0.3, alpha=0.2, eps=1e-08, mix=gaussian)"


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=lambda c: f"{c['idx']}-{c['mix']}")
def test_cuda_module_matches_reference_mixstyle(golden, c, monkeypatch):
    from maxstyle_b200 import MixStyle
    x_np, dy_np, rec = _case(golden, c)
    m = MixStyle(p=1.0, alpha=0.1, mix=c["mix"], lmda=c["fixed_lmda"])
    if c["sampler"] is not None:
        m.coeficient_sampler = c["sampler"]
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    if c["mix"] == "gaussian":
        # the two device draws come from the CUDA generator here (CPU in the fixture): feed the fixture's values
        draws = [torch.from_numpy(rec["eps_mu"]).cuda().view(*rec["eps_mu"].shape, 1, 1),
                 torch.from_numpy(rec["eps_sig"]).cuda().view(*rec["eps_sig"].shape, 1, 1)]
        real = torch.randn

        def fake_randn(*a, **k):
            if k.get("device") is not None and torch.device(k["device"]).type == "cuda":
                return draws.pop(0)
            return real(*a, **k)
        monkeypatch.setattr(torch, "randn", fake_randn)
    torch.manual_seed(c["seed"])
    given = torch.tensor([1, 0, 3, 2, 5, 4]) if c["given_perm"] else None
    y = m(x, perm=given) if given is not None else m(x)
    if c["mix"] != "gaussian":
        assert float(torch.rand(1)) == c["next_cpu_rand"], "CPU generator consumed differently from the reference"
        assert np.array_equal(np.asarray(m.get_perm()), rec["perm"]), "permutation must be bit-exact"
    y.backward(torch.from_numpy(dy_np).cuda())
    tol_y, tol_g = (5e-3, 5e-3) if c["kind"] == "offset" else (1e-5, 1e-4)
    assert _rel(y.detach().cpu().numpy(), rec["y"]) < tol_y
    assert _rel(x.grad.cpu().numpy(), rec["dx"]) < tol_g
    # and against the float64 oracle, which for 'offset' inputs is the tighter check
    y64, cache = MO.mixstyle_forward(x_np, c["mix"], lmda=rec.get("lmda"), perm=rec.get("perm"), eps_mu=rec.get("eps_mu"),
                                     eps_sig=rec.get("eps_sig"), eps=1e-8, dtype=np.float64)
    assert _rel(y.detach().cpu().numpy(), y64) < 1e-5
    assert _rel(x.grad.cpu().numpy(), MO.mixstyle_backward(dy_np, cache, dtype=np.float64)) < 1e-5


@pytest.mark.gpu
def test_cuda_mixstyle_channels_last_and_bf16():
    from maxstyle_b200 import MixStyle
    torch.manual_seed(5)
    x = (torch.randn(8, 16, 24, 24, device="cuda") * 1.5 + 0.3)
    m = MixStyle(p=1.0, mix="random")
    torch.manual_seed(9)
    y0 = m(x)
    torch.manual_seed(9)
    y1 = m(x.contiguous(memory_format=torch.channels_last))
    assert y1.is_contiguous(memory_format=torch.channels_last)
    assert float((y0 - y1).abs().max()) < 1e-5 * float(y0.abs().max())
    torch.manual_seed(9)
    y2 = m(x.bfloat16())
    assert y2.dtype == torch.bfloat16 and float((y0 - y2.float()).abs().max()) < 2 ** -6 * float(y0.abs().max())
