"""numpy restatement of faiss `IndexIVFFlat(IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT)`
as the reference builds and uses it (faiss_retrieval.py:50-55, :93, :118, :150-155).  TEST INFRASTRUCTURE.

Semantics restated from faiss 1.7.x (PARITY UNPINNED against faiss itself, see oracle/__init__.py):
  train  : k-means, niter = 10, at most 256 training points per centroid (seeded subsample, seed 1234),
           spherical for the inner-product metric (centroids L2-normalised after every update),
           assignment by the IndexFlatIP quantiser = maximum inner product
  add    : list = argmax_c <x, centroid_c>; the vector is appended to that inverted list
  search : the `nprobe` centroids of largest inner product, exact fp32 inner products over the
           vectors of those lists, top-k best-first; unfilled slots are (-FLT_MAX, -1)
Because faiss's RNG / subsampling cannot be reproduced bit-for-bit without its source, IVF parity
is defined on SHARED centroids: `set_centroids()` injects the GPU index's (or any) centroids, after
which list membership, probed lists and results must be identical.
"""
from __future__ import annotations

import numpy as np

from .flat import NEG_FLT_MAX, normalize_L2


def kmeans(x: np.ndarray, k: int, niter: int = 10, seed: int = 1234, spherical: bool = True,
           max_points_per_centroid: int = 256) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    if n < k:
        raise ValueError(f"fewer training vectors ({n}) than centroids ({k})")
    rng = np.random.default_rng(seed)
    if n > k * max_points_per_centroid:
        x = x[rng.permutation(n)[:k * max_points_per_centroid]]
        n = x.shape[0]
    cent = x[rng.permutation(n)[:k]].copy()
    if spherical:
        normalize_L2(cent)
    for it in range(niter):
        assign = np.argmax(x @ cent.T, axis=1)
        sums = np.zeros_like(cent)
        np.add.at(sums, assign, x)
        counts = np.bincount(assign, minlength=k)
        empty = counts == 0
        cent = sums / np.maximum(counts, 1)[:, None].astype(np.float32)
        if empty.any():
            cent[empty] = x[rng.integers(0, n, int(empty.sum()))]
        if spherical:
            normalize_L2(cent)
    return cent.astype(np.float32)


def assign_max_ip(x: np.ndarray, centroids: np.ndarray, chunk: int = 1 << 16) -> np.ndarray:
    """argmax_c <x, centroid_c> per row (the IndexFlatIP quantiser), computed chunk by chunk so that a
    10M-row corpus never materialises its [n, nlist] score matrix; ties -> lowest list id (np.argmax)."""
    out = np.empty(len(x), dtype=np.int64)
    ct = np.ascontiguousarray(centroids.T)
    for lo in range(0, len(x), chunk):
        out[lo:lo + chunk] = np.argmax(x[lo:lo + chunk] @ ct, axis=1)
    return out


def rows_by_list(assign: np.ndarray, nlist: int):
    """(order, offsets): rows of list l, ascending, are order[offsets[l]:offsets[l+1]] — the same row sets
    as `np.nonzero(assign == l)`, found once instead of once per (query, list)."""
    order = np.argsort(assign, kind="stable")
    offsets = np.zeros(nlist + 1, dtype=np.int64)
    np.cumsum(np.bincount(assign, minlength=nlist), out=offsets[1:])
    return order, offsets


class OracleIndexIVFFlat:
    def __init__(self, d: int, nlist: int, **_):
        self.d, self.nlist = d, nlist
        self._by_list = None
        self.nprobe = 1
        self.centroids = None
        self.xb = np.zeros((0, d), dtype=np.float32)
        self.assign = np.zeros(0, dtype=np.int64)

    @property
    def is_trained(self) -> bool:
        return self.centroids is not None

    @property
    def ntotal(self) -> int:
        return self.xb.shape[0]

    def train(self, x) -> None:
        if not self.is_trained:
            self.centroids = kmeans(np.asarray(x, dtype=np.float32), self.nlist)

    def set_centroids(self, c) -> None:
        c = np.ascontiguousarray(c, dtype=np.float32)
        assert c.shape == (self.nlist, self.d)
        self.centroids = c

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        self.xb = np.concatenate([self.xb, x])
        self.assign = np.concatenate([self.assign, assign_max_ip(x, self.centroids)])
        self._by_list = None

    def list_sizes(self) -> np.ndarray:
        return np.bincount(self.assign, minlength=self.nlist).astype(np.int64)

    def probe(self, q) -> np.ndarray:
        """[Q, nprobe] list ids, best centroid first (ties by list id)."""
        S = np.ascontiguousarray(q, dtype=np.float32) @ self.centroids.T
        npb = min(self.nprobe, self.nlist)
        return np.stack([np.lexsort((np.arange(self.nlist), -row))[:npb] for row in S])

    def search(self, q, k: int, extra: int = 0):
        q = np.ascontiguousarray(q, dtype=np.float32)
        kk = k + extra
        D = np.full((len(q), kk), NEG_FLT_MAX, dtype=np.float32)
        I = np.full((len(q), kk), -1, dtype=np.int64)
        lists = self.probe(q)
        if self._by_list is None or self._by_list[2] != len(self.assign):
            self._by_list = rows_by_list(self.assign, self.nlist) + (len(self.assign),)
        order_l, off = self._by_list[:2]
        for qi in range(len(q)):
            rows = np.sort(np.concatenate([order_l[off[l]:off[l + 1]] for l in lists[qi]]))
            if rows.size == 0:
                continue
            s = self.xb[rows] @ q[qi]
            order = np.lexsort((rows, -s))[:kk]
            D[qi, :order.size] = s[order]
            I[qi, :order.size] = rows[order]
        return D, I
