"""Global-batch MaxStyle across the GPUs of one node (BASELINE.json: "the batch shards
data-parallel ... the only collective on this path is an all-gather of the NxC mu/sigma tables,
so mixing partners can be drawn from the global batch").  No reference counterpart: the
reference is single-GPU; the oracle for an R-rank run is the reference on the concatenated batch.

Rank r owns samples [r*N_local, (r+1)*N_local) of a global batch of R*N_local and the matching
rows of gamma_noise / beta_noise / lmda (parameters are per-sample, so there is no gradient
all-reduce).  `perm` is a permutation of the GLOBAL batch, identical on every rank.  Forward:
stats kernel -> one all-gather of the packed [N_local, 2C] (mu | sig) rows -> table kernel over
the global table -> apply kernel.  Backward needs no collective (partner statistics are detached).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib as L
from . import functional as F
from .layer import MaxStyle


class StyleTableExchange:
    """Host-side plumbing of the one collective on the path: row partition + all-gather of the
    packed style table.  Backend-agnostic (NCCL on GPUs; gloo on CPU tensors in the unit tests)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("StyleTableExchange needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def rows(self, n_local: int):
        """(row_offset, n_global) of this rank's slab; every rank holds the same n_local."""
        return self.rank * n_local, self.world * n_local

    def allocate(self, n_local: int, c: int, device) -> torch.Tensor:
        """Packed global table [N_global, 2C]: columns [0,C) are mu, [C,2C) are sig."""
        return torch.empty(self.world * n_local, 2 * c, dtype=torch.float32, device=device)

    @staticmethod
    def views(table: torch.Tensor):
        c = table.shape[1] // 2
        return table[:, :c], table[:, c:]

    def gather(self, table: torch.Tensor, n_local: int) -> torch.Tensor:
        """All-gather the local rows of `table` into every rank's copy (in place)."""
        off, _ = self.rows(n_local)
        mine = table[off:off + n_local]
        if table.is_cuda:
            dist.all_gather_into_tensor(table, mine, group=self.group)      # NCCL in-place all-gather
        else:
            dist.all_gather_into_tensor(table, mine.clone(), group=self.group)
        return table

    def agree(self, t: torch.Tensor, device=None) -> torch.Tensor:
        """Broadcast a small CPU tensor (perm, rand_p) from rank 0 so that every rank uses rank 0's draw."""
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            buf = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
            dist.broadcast(buf, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            return buf.cpu()
        buf = t.clone()
        dist.broadcast(buf, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        return buf


class PeerTableExchange:
    """The exchange buffers of `maxstyle_tables_p2p` (include/maxstyle_b200.h): one small buffer per rank in symmetric
    memory (torch.distributed._symmetric_memory: allocated on every rank, mapped into every peer over NVLink), the device
    array of the peers' addresses, and the device-side epoch / done / error words.  Construction is collective (every rank
    of the group must call it, in the same order).  Raises if symmetric memory is not available -- callers that can also
    use the NCCL all-gather (GraphedLayerStep) decide what to do then."""

    def __init__(self, n_local: int, c: int, device, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        nbytes = int(L.get_lib().maxstyle_p2p_bytes(n_local, c, self.world))
        self.buffer = symm.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.buffer.zero_()
        self.handle = symm.rendezvous(self.buffer, group=group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.buffer.data_ptr():
            raise RuntimeError("maxstyle_b200: symmetric-memory rendezvous returned unexpected peer pointers")
        self.peers_dev = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.n_local, self.channels = n_local, c
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.bar_epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self.handle.barrier()                        # every rank's buffer is zeroed before anyone publishes into the scheme

    def check(self):
        """Synchronises; raises if a wait for a peer timed out since construction (debug / tests)."""
        if int(self.error.item()) != 0:
            raise RuntimeError("maxstyle_b200: maxstyle_tables_p2p timed out waiting for a peer rank; results are invalid")


class GlobalBatchMaxStyle(MaxStyle):
    """MaxStyle whose mixing partners and batch statistics span the global data-parallel batch.

    `batch_size` is the LOCAL batch (what this rank's feature map has); the random state is drawn
    for the global batch exactly like the reference would on the concatenated batch -- seed every
    rank identically (torch.manual_seed(s)) to reproduce the single-device reference bit for bit.
    With different seeds per rank the layer is still consistent: perm and rand_p are rank 0's on
    every rank (agreed before anything is allocated from them), each rank's parameter rows come
    from its own draw -- and this rank keeps its rows.
    """

    def __init__(self, batch_size, num_feature, p=0.5, mix_style=True, no_noise=False, mix_learnable=True,
                 noise_learnable=True, always_use_beta=False, alpha=0.1, eps=1e-6, use_gpu=True, debug=False,
                 *, group: Optional[dist.ProcessGroup] = None):
        self._exchange = StyleTableExchange(group)
        self.local_batch_size = batch_size
        self.row_offset, self.global_batch_size = self._exchange.rows(batch_size)
        self._table = None
        super().__init__(batch_size, num_feature, p=p, mix_style=mix_style, no_noise=no_noise,
                         mix_learnable=mix_learnable, noise_learnable=noise_learnable, always_use_beta=always_use_beta,
                         alpha=alpha, eps=eps, use_gpu=use_gpu, debug=debug)

    def _agree_on_draw(self):
        """Every rank takes rank 0's permutation and rank 0's activity draw -- BEFORE the parameters are allocated from
        rand_p (ranks seeded differently would otherwise disagree on whether the layer is active, and a rank whose own
        draw was inactive would run an active forward with zero, non-learnable parameters)."""
        dev = self.device if self.use_gpu else None
        self.perm = self._exchange.agree(self.perm, dev)
        self.rand_p = self._exchange.agree(self.rand_p, dev)

    def init_parameters(self):
        n_loc, n_glob, off = self.local_batch_size, self.global_batch_size, self.row_offset
        self.batch_size = n_glob                       # draw the reference's state for the global batch ...
        try:
            super().init_parameters()
        finally:
            self.batch_size = n_loc
        self._perm_dev = None

        def local_rows(t):                             # ... and keep this rank's rows
            rows = t.detach()[off:off + n_loc].clone()
            if isinstance(t, nn.Parameter):
                return nn.Parameter(rows, requires_grad=t.requires_grad)
            rows.requires_grad = False
            return rows

        self.gamma_noise = local_rows(self.gamma_noise)
        self.beta_noise = local_rows(self.beta_noise)
        self.lmda = local_rows(self.lmda)

    def _stat_table(self, device) -> torch.Tensor:
        if self._table is None or self._table.device != device:
            self._table = self._exchange.allocate(self.local_batch_size, self.num_feature, device)
        return self._table

    def forward(self, x):
        self.data = x
        n, c = x.size(0), x.size(1)
        plane = x.numel() // (n * c) if n * c else 0
        # identity cases; note B <= 1 refers to the GLOBAL batch here
        if (self.rand_p >= self.p) or (not self.mix_style and self.no_noise) or self.global_batch_size <= 1 or plane == 1:
            return x
        assert self.local_batch_size == n and self.num_feature == c, \
            f"check input dim, expect ({self.local_batch_size}, {self.num_feature}, *,*) , got {n}{c}"
        if not x.is_cuda:
            raise RuntimeError("maxstyle_b200: GlobalBatchMaxStyle.forward got a CPU tensor; there is no CPU fallback")
        with torch.cuda.device(x.device):
            return GlobalBatchFunction.apply(x, self.gamma_noise, self.beta_noise, self.lmda, self)


class GlobalBatchFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma_noise, beta_noise, lmda, layer):
        x = F.dense_layout(x)
        n, c = x.shape[0], x.shape[1]
        flags = layer._flags()
        first = layer.gamma_std is None or layer.beta_std is None
        if first:
            gamma_std = torch.empty(1, c, 1, 1, dtype=torch.float32, device=x.device)
            beta_std = torch.empty(1, c, 1, 1, dtype=torch.float32, device=x.device)
            flags |= L.FLAG_COMPUTE_BATCH_STD
        else:
            gamma_std, beta_std = layer.gamma_std, layer.beta_std
        ws = layer._workspace_for(x)
        # a fresh table per forward: backward of an earlier forward may still need its statistics
        table = layer._exchange.allocate(n, c, x.device)
        mu_all, sig_all = StyleTableExchange.views(table)
        F.instance_stats(x, layer.eps, ws, mu_all, sig_all, layer.row_offset)
        layer._exchange.gather(table, n)                                   # the one collective on the path
        perm_dev = layer._perm_device(x.device)
        scale, shift = F.style_tables(mu_all, sig_all, layer.row_offset, n, perm_dev, lmda, gamma_noise, beta_noise,
                                      gamma_std, beta_std, flags)
        y = F.style_apply(x, mu_all, layer.row_offset, scale, shift)
        if first:
            layer.gamma_std, layer.beta_std = gamma_std, beta_std
        ctx.layer = layer
        ctx.flags = flags & ~L.FLAG_COMPUTE_BATCH_STD
        ctx.tables = (mu_all, sig_all, scale, gamma_std, beta_std)
        ctx.save_for_backward(x, lmda)
        ctx.set_materialize_grads(False)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        if dy is None:
            return None, None, None, None, None
        x, lmda = ctx.saved_tensors
        layer = ctx.layer
        mu_all, sig_all, scale, gamma_std, beta_std = ctx.tables
        need_dx, need_g, need_b, need_l = ctx.needs_input_grad[:4]
        fused = layer._fused_step
        step = fused.struct(layer.gamma_noise, layer.beta_noise, layer.lmda) if fused is not None else None
        keep = fused is None or fused.keep_grads
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dy = F._match_layout(dy, x)
        with torch.cuda.device(x.device):
            dx, dg, db, dl = F.backward_raw(dy, x, mu_all, sig_all, layer.row_offset, scale,
                                            layer._perm_device(x.device), lmda, gamma_std, beta_std, ctx.flags,
                                            layer._workspace_for(x), need_dx=need_dx,
                                            need_noise_grad=(need_g or need_b) and keep,
                                            need_mix_grad=need_l and keep, step=step)
        n, c = x.shape[0], x.shape[1]
        if not keep:
            return dx, None, None, None, None
        return (dx,
                dg.view(n, c, 1, 1) if (need_g and dg is not None) else None,
                db.view(n, c, 1, 1) if (need_b and db is not None) else None,
                dl.view(n, 1, 1, 1) if (need_l and dl is not None) else None,
                None)
