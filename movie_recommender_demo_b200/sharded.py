"""Row-sharded Flat index over several B200s (the reference has no sharding —
one faiss.Index object, faiss_retrieval.py:42-78).

One process per GPU.  Rank r owns the contiguous corpus rows [r*N/P, (r+1)*N/P) and returns
global labels (local row + base).  Every rank searches ITS shard for the SAME query batch;
the per-rank best-first top-k lists (scores fp32 + labels int64, Q*k*12 bytes per rank) are
exchanged with ONE all-gather (NCCL over NVLink/NVSwitch) and merged by `b2r_topk_merge` on
every rank.  Because each shard's scores are exact fp32 (rescored) before the exchange, the
merged result is identical to an unsharded search.
"""
from __future__ import annotations

from typing import Tuple

from . import _lib

__all__ = ["shard_rows", "gather_topk", "ShardedFlatIndex"]


def shard_rows(total_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; ranges tile [0, total_rows) exactly."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return rank * total_rows // world, (rank + 1) * total_rows // world


def gather_topk(D_local, I_local, group=None):
    """All-gather the per-rank [Q,k] results -> ([P,Q,k] scores, [P,Q,k] labels), rank-major."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    D_all = torch.empty((world,) + tuple(D_local.shape), dtype=D_local.dtype, device=D_local.device)
    I_all = torch.empty((world,) + tuple(I_local.shape), dtype=I_local.dtype, device=I_local.device)
    if D_local.is_cuda:
        dist.all_gather_into_tensor(D_all, D_local.contiguous(), group=group)
        dist.all_gather_into_tensor(I_all, I_local.contiguous(), group=group)
    else:  # gloo (host-logic tests)
        dist.all_gather(list(D_all.unbind(0)), D_local.contiguous(), group=group)
        dist.all_gather(list(I_all.unbind(0)), I_local.contiguous(), group=group)
    return D_all, I_all


class ShardedFlatIndex:
    """Flat inner-product index whose rows are sharded over the ranks of a process group."""

    def __init__(self, d: int, total_rows: int, group=None, device=None):
        import torch
        import torch.distributed as dist
        from .faiss_retrieval import IndexFlatIP
        self.d = d
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total_rows = total_rows
        self.lo, self.hi = shard_rows(total_rows, self.world, self.rank)
        self.local = IndexFlatIP(d, device=device)
        self.local.set_label_base(self.lo)
        self._torch = torch

    @property
    def ntotal_local(self) -> int:
        return self.local.ntotal

    def add_local(self, x, normalize: bool = True) -> None:
        """Append rows of THIS rank's range (callers feed lo..hi in order)."""
        if self.local.ntotal + len(x) > self.hi - self.lo:
            raise ValueError("more rows than this rank's shard holds")
        self.local.add(x, normalize=normalize)

    def search_device(self, q, k: int, normalize: bool = True):
        """(D [Q,k], I [Q,k], status) CUDA tensors, identical on every rank."""
        torch = self._torch
        Dl, Il, st, _ = self.local.search_device(q, k, normalize=normalize)
        if self.world == 1:
            return Dl, Il, st
        D_all, I_all = gather_topk(Dl, Il, self.group)
        Q = Dl.shape[0]
        D_out = torch.empty_like(Dl)
        I_out = torch.empty_like(Il)
        lib = _lib.load()
        _lib.check(lib.b2r_topk_merge(self.world, Q, k, D_all.data_ptr(), I_all.data_ptr(), D_out.data_ptr(),
                                      I_out.data_ptr(), 1,
                                      int(torch.cuda.current_stream(Dl.device).cuda_stream)))
        return D_out, I_out, st
