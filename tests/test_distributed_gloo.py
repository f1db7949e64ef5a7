"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row partition, perm agreement,
parameter slabs and the packed-table all-gather.  The kernels cannot run here, so the per-rank
compute is done by the oracle -- which is exactly the claim being tested: R ranks on row slabs
+ one all-gather of the [N,2C] tables == the reference on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        import json
        from maxstyle_b200 import GlobalBatchMaxStyle, MaxStyle, StyleTableExchange
        from oracle import maxstyle_oracle as O
        from oracle.gen_golden import make_input
        torch.set_num_threads(1)
        man = json.load(open(os.path.join(ROOT, "tests", "golden", "MANIFEST.json")))
        idx = [m["name"] for m in man["fwd_bwd"]].index("mid_16ch")
        meta = man["fwd_bwd"][idx]
        g = np.load(os.path.join(ROOT, "tests", "golden", "fwd_bwd.npz"))
        pre = f"f{idx}_"
        n_glob, c, h, w = meta["N"], meta["C"], meta["H"], meta["W"]
        n_loc = n_glob // WORLD

        # ---- 1. state: same seed on every rank == the single-device module, sliced ----------
        torch.manual_seed(meta["seed"] + rank * 0)
        layer = GlobalBatchMaxStyle(n_loc, c, p=1.0, use_gpu=False)
        torch.manual_seed(meta["seed"])
        single = MaxStyle(n_glob, c, p=1.0, use_gpu=False)
        off = rank * n_loc
        assert layer.row_offset == off and layer.global_batch_size == n_glob and layer.batch_size == n_loc
        assert torch.equal(layer.perm, single.perm) and torch.equal(layer.rand_p, single.rand_p)
        assert np.array_equal(layer.perm.numpy(), g[pre + "perm"])            # == the reference's draw
        for k in ("gamma_noise", "beta_noise", "lmda"):
            assert torch.equal(getattr(layer, k).data, getattr(single, k).data[off:off + n_loc]), k
            assert isinstance(getattr(layer, k), torch.nn.Parameter)
        assert [k for k, _ in layer.named_parameters()] == ["gamma_noise", "beta_noise", "lmda"]

        # ---- 2. perm agreement even when ranks were seeded differently ----------------------
        torch.manual_seed(1000 + rank)
        other = GlobalBatchMaxStyle(n_loc, c, p=1.0, use_gpu=False)
        gathered = [torch.empty_like(other.perm) for _ in range(WORLD)]
        dist.all_gather(gathered, other.perm)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        assert sorted(other.perm.tolist()) == list(range(n_glob))

        # ---- 3. table exchange + sharded compute == reference on the concatenated batch ----
        x = make_input(meta["seed"], (n_glob, c, h, w), meta["kind"])
        dy = np.random.RandomState(meta["seed"] + 5000).standard_normal(size=(n_glob, c, h, w)).astype(np.float32)
        xs, dys = x[off:off + n_loc], dy[off:off + n_loc]
        ex = layer._exchange
        table = ex.allocate(n_loc, c, torch.device("cpu"))
        table.fill_(float("nan"))
        mu_all, sig_all = StyleTableExchange.views(table)
        mu_l, sig_l = O.instance_stats(xs, 1e-6)
        mu_all[off:off + n_loc] = torch.from_numpy(mu_l)
        sig_all[off:off + n_loc] = torch.from_numpy(sig_l)
        ex.gather(table, n_loc)
        assert not torch.isnan(table).any() and table.shape == (n_glob, 2 * c)
        assert np.allclose(mu_all.numpy(), g[pre + "mu"], rtol=1e-5, atol=1e-6)
        st = O.StyleState(perm=layer.perm.numpy(), gamma_noise=layer.gamma_noise.detach().numpy().reshape(n_loc, c),
                          beta_noise=layer.beta_noise.detach().numpy().reshape(n_loc, c),
                          lmda=layer.lmda.detach().numpy().reshape(n_loc), p=1.0)
        y, cache = O.forward(xs, st, global_mu=mu_all.numpy(), global_sig=sig_all.numpy(), row_offset=off)
        dx, dgam, dbet, dlm = O.backward(dys, xs, st, cache, global_mu=mu_all.numpy(), global_sig=sig_all.numpy(),
                                         row_offset=off)

        def close(a, b, rtol):
            return np.abs(a - b).max() <= rtol * np.abs(b).max()

        assert close(y, g[pre + "y"][off:off + n_loc], 1e-5)
        assert close(dx, g[pre + "dx"][off:off + n_loc], 1e-4)
        assert close(dgam, g[pre + "d_gamma_noise"][off:off + n_loc], 1e-4)
        assert close(dbet, g[pre + "d_beta_noise"][off:off + n_loc], 1e-4)
        assert np.abs(dlm - g[pre + "d_lmda"].reshape(-1)[off:off + n_loc]).max() <= 1e-4 * np.abs(g[pre + "d_lmda"]).max()
        # CPU tensors never reach the kernels: loud failure, no fallback
        try:
            layer(torch.from_numpy(xs))
            raise SystemExit("expected RuntimeError")
        except RuntimeError as e:
            assert "no CPU fallback" in str(e)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_global_batch(tmp_path):
    mp.spawn(_worker, args=(_free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(WORLD))
