"""numpy restatement of faiss `IndexIVFPQ(IndexFlatIP(d), d, nlist, m, 8)` as the reference builds it
(faiss_retrieval.py:57-63: no metric argument => METRIC_L2, by_residual = True).  TEST INFRASTRUCTURE.
PARITY UNPINNED against faiss itself (see oracle/__init__.py); parity with the GPU index is defined on
SHARED centroids and codebooks.

  train  : coarse k-means (not spherical: the metric is L2; assignment still by the IndexFlatIP
           quantiser = max inner product), then per sub-space 256-means on the residuals
           (25 iterations, <= 65536 training points)
  add    : residual r = x - c(list); code_s = argmin_j |r_s - codeword_{s,j}|^2
  search : top-nprobe centroids by inner product; ADC distance sum_s |(q - c)_s - codeword_{s,code_s}|^2
           over the probed lists; k SMALLEST distances, ascending; unfilled slots (+FLT_MAX, -1)
"""
from __future__ import annotations

import numpy as np

from .ivf import assign_max_ip, kmeans, rows_by_list

FLT_MAX = np.float32(3.4028234663852886e38)


def _l2_kmeans(x, k, niter, seed):
    rng = np.random.default_rng(seed)
    cent = x[rng.permutation(len(x))[:k]].copy()
    for _ in range(niter):
        d2 = (x * x).sum(1)[:, None] - 2 * x @ cent.T + (cent * cent).sum(1)[None]
        a = d2.argmin(1)
        sums = np.zeros_like(cent)
        np.add.at(sums, a, x)
        cnt = np.bincount(a, minlength=k)
        cent = sums / np.maximum(cnt, 1)[:, None].astype(np.float32)
        if (cnt == 0).any():
            cent[cnt == 0] = x[rng.integers(0, len(x), int((cnt == 0).sum()))]
    return cent.astype(np.float32)


class OracleIndexIVFPQ:
    def __init__(self, d, nlist, m=8, nbits=8, **_):
        assert nbits == 8 and d % m == 0
        self.d, self.nlist, self.m, self.dsub = d, nlist, m, d // m
        self.nprobe = 1
        self.centroids = None
        self.codebooks = None          # [m, 256, dsub]
        self.codes = np.zeros((0, m), dtype=np.uint8)
        self.assign = np.zeros(0, dtype=np.int64)

    @property
    def is_trained(self):
        return self.centroids is not None and self.codebooks is not None

    @property
    def ntotal(self):
        return len(self.codes)

    def train(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if self.centroids is None:
            self.centroids = kmeans(x, self.nlist, spherical=False)
        if self.codebooks is None:
            rng = np.random.default_rng(1234)
            xt = x[rng.permutation(len(x))[:65536]]
            r = xt - self.centroids[np.argmax(xt @ self.centroids.T, axis=1)]
            self.codebooks = np.stack([_l2_kmeans(np.ascontiguousarray(r[:, s * self.dsub:(s + 1) * self.dsub]), 256, 25, s)
                                       for s in range(self.m)])

    def set_centroids(self, c):
        self.centroids = np.ascontiguousarray(c, dtype=np.float32)

    def set_codebooks(self, cb):
        cb = np.ascontiguousarray(cb, dtype=np.float32)
        assert cb.shape == (self.m, 256, self.dsub)
        self.codebooks = cb

    def encode(self, x, assign):
        r = x - self.centroids[assign]
        codes = np.empty((len(x), self.m), dtype=np.uint8)
        for s in range(self.m):
            rs = r[:, s * self.dsub:(s + 1) * self.dsub]
            cb = self.codebooks[s]
            d2 = (rs * rs).sum(1)[:, None] - 2 * rs @ cb.T + (cb * cb).sum(1)[None]
            codes[:, s] = d2.argmin(1)
        return codes

    def add(self, x, codes=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        a = assign_max_ip(x, self.centroids)
        c = self.encode(x, a) if codes is None else np.ascontiguousarray(codes, dtype=np.uint8)
        self.assign = np.concatenate([self.assign, a])
        self.codes = np.concatenate([self.codes, c])

    def list_sizes(self):
        return np.bincount(self.assign, minlength=self.nlist).astype(np.int64)

    def search(self, q, k, extra=0):
        q = np.ascontiguousarray(q, dtype=np.float32)
        kk = k + extra
        D = np.full((len(q), kk), FLT_MAX, dtype=np.float32)
        I = np.full((len(q), kk), -1, dtype=np.int64)
        S = q @ self.centroids.T
        npb = min(self.nprobe, self.nlist)
        order_l, off = rows_by_list(self.assign, self.nlist)
        for qi in range(len(q)):
            lists = np.lexsort((np.arange(self.nlist), -S[qi]))[:npb]
            rows_all, dist_all = [], []
            for l in lists:
                rows = order_l[off[l]:off[l + 1]]
                if rows.size == 0:
                    continue
                res = (q[qi] - self.centroids[l]).reshape(self.m, self.dsub)
                lut = ((res[:, None, :] - self.codebooks) ** 2).sum(-1).astype(np.float32)       # [m, 256]
                dist = lut[np.arange(self.m)[None, :], self.codes[rows]].sum(1, dtype=np.float32)
                rows_all.append(rows)
                dist_all.append(dist)
            if not rows_all:
                continue
            rows = np.concatenate(rows_all)
            dist = np.concatenate(dist_all)
            order = np.lexsort((rows, dist))[:kk]
            D[qi, :order.size] = dist[order]
            I[qi, :order.size] = rows[order]
        return D, I
