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

__all__ = ["shard_rows", "gather_topk", "ShardedFlatIndex", "ShardedIVFIndex"]


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

    _largest = 1   # inner product: larger is better (b2r_topk_merge `largest`)

    def __init__(self, d: int, total_rows: int, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.d = d
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.total_rows = total_rows
        self.lo, self.hi = shard_rows(total_rows, self.world, self.rank)
        self.local = self._make_local(device)
        self.local.set_label_base(self.lo)
        self.local.reserve(self.hi - self.lo)      # one allocation for the whole shard (no growth copies)
        self._torch = torch

    def _make_local(self, device):
        from .faiss_retrieval import IndexFlatIP
        return IndexFlatIP(self.d, device=device)

    @property
    def ntotal_local(self) -> int:
        return self.local.ntotal

    def add_local(self, x, normalize: bool = True) -> None:
        """Append rows of THIS rank's range (callers feed lo..hi in order)."""
        if self.local.ntotal + len(x) > self.hi - self.lo:
            raise ValueError("more rows than this rank's shard holds")
        self.local.add(x, normalize=normalize)

    def search_device(self, q, k: int, normalize: bool = True, **local_kw):
        """(D [Q,k], I [Q,k], status) CUDA tensors, identical on every rank."""
        torch = self._torch
        Dl, Il, st, _ = self.local.search_device(q, k, normalize=normalize, **local_kw)
        if self.world == 1:
            return Dl, Il, st
        D_all, I_all = gather_topk(Dl, Il, self.group)
        Q = Dl.shape[0]
        D_out = torch.empty_like(Dl)
        I_out = torch.empty_like(Il)
        lib = _lib.load()
        _lib.check(lib.b2r_topk_merge(self.world, Q, k, D_all.data_ptr(), I_all.data_ptr(), D_out.data_ptr(),
                                      I_out.data_ptr(), self._largest,
                                      int(torch.cuda.current_stream(Dl.device).cuda_stream)))
        return D_out, I_out, st


class ShardedIVFIndex(ShardedFlatIndex):
    """IVF-Flat (inner product) or IVF-PQ (L2) index sharded by row ownership: the coarse
    centroids (and PQ codebooks) are REPLICATED, every rank files its own rows into its own copy of the
    inverted lists, every rank probes the same `nprobe` lists for the same queries, and the per-rank
    top-k lists take the same all-gather + merge as the flat index.  With shared quantisers the merged
    answer equals an unsharded index's: IVF-Flat scores are exact fp32 before the exchange, IVF-PQ ADC
    scores depend only on (query, list, code)."""

    def __init__(self, d: int, total_rows: int, nlist: int, nprobe: int = 10, kind: str = 'IVF', pq_m: int = 8,
                 group=None, device=None):
        if kind not in ('IVF', 'IVFPQ'):
            raise ValueError(f"Unknown index type: {kind}")
        self.kind, self.nlist, self.nprobe, self.pq_m = kind, int(nlist), int(nprobe), int(pq_m)
        self._largest = 1 if kind == 'IVF' else 0          # IVF-PQ is L2: ascending distances
        super().__init__(d, total_rows, group=group, device=device)

    def _make_local(self, device):
        from . import ivf
        if self.kind == 'IVF':
            return ivf.IndexIVFFlat(self.d, self.nlist, device=device)
        return ivf.IndexIVFPQ(self.d, self.nlist, self.pq_m, device=device)

    def train(self, x, src: int = 0) -> None:
        """Collective.  Rank `src` trains on ITS `x` (other ranks' `x` is ignored and may be None); the
        centroids / codebooks are then broadcast so that every rank holds bit-identical quantisers."""
        import torch.distributed as dist
        torch = self._torch
        dev = self.local.device
        if self.rank == src:
            self.local.train(x)
        cent = torch.empty((self.nlist, self.d), dtype=torch.float32, device=dev)
        cb = torch.empty((self.pq_m, 256, self.d // self.pq_m), dtype=torch.float32, device=dev) \
            if self.kind == 'IVFPQ' else None
        if self.rank == src:
            cent.copy_(torch.from_numpy(self.local.export_centroids()))
            if cb is not None:
                cb.copy_(torch.from_numpy(self.local.export_codebooks()))
        if self.world > 1:
            dist.broadcast(cent, src=src, group=self.group)
            if cb is not None:
                dist.broadcast(cb, src=src, group=self.group)
        if self.rank != src:
            self.import_quantisers(cent.cpu().numpy(), cb.cpu().numpy() if cb is not None else None)

    def import_quantisers(self, centroids, codebooks=None) -> None:
        if self.kind == 'IVFPQ':
            if codebooks is None:
                raise ValueError("IVFPQ needs codebooks")
            self.local.import_codebooks(codebooks)
        self.local.import_centroids(centroids)

    def add_local(self, x, normalize: bool = True) -> None:
        if not self.local.is_trained:
            raise RuntimeError("train() (collective) or import_quantisers() first: every rank must share the "
                               "same quantisers before rows are filed")
        super().add_local(x, normalize=normalize)

    def search_device(self, q, k: int, normalize: bool = True, nprobe: int = 0):
        return super().search_device(q, k, normalize=normalize, nprobe=nprobe or self.nprobe)
