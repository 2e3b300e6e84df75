"""Exact fp32 restatement of the reference's Flat inner-product retrieval (numpy).

Follows, in order of operations:
  * `faiss.normalize_L2`                      faiss_retrieval.py:115, :147
        per row: if sum(x^2) > 0: x *= 1/sqrt(sum(x^2))   (zero rows stay zero), fp32
  * `faiss.IndexFlatIP.add / .search`         faiss_retrieval.py:48, :118, :155
        D = q @ X.T in fp32, top-k best-first (descending inner product), labels int64,
        missing slots label -1 / distance -3.4028235e38
  * `FAISSIndex` wrapper semantics            faiss_retrieval.py:97-127, :129-166, :168-194, :247-256
        astype('float32') copies (inputs never mutated), default ids continue from
        len(id_map), id remap `id_map[idx]` with python negative-index wrap for -1,
        return order (ad_ids, distances).

  * `faiss.IndexHNSWFlat(d, 32)`               faiss_retrieval.py:65-70
        default metric L2, distances ascending.  HNSW is an APPROXIMATION of the exact L2 nearest
        neighbours whose graph depends on faiss's RNG and insertion threading, so it has no reproducible
        answer of its own; its oracle is the thing it approximates: exact fp32 ||q - x||^2, top-k smallest
        (`OracleIndexFlatL2`), missing slots label -1 / distance +3.4028235e38.

faiss's order among exactly equal scores is implementation-defined; the oracle's canonical
order is (-score, label).  PARITY UNPINNED against faiss itself (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

NEG_FLT_MAX = np.float32(-3.4028234663852886e38)


def normalize_L2(x: np.ndarray) -> np.ndarray:
    """In-place row normalisation of a C-contiguous float32 array (faiss.normalize_L2)."""
    assert x.dtype == np.float32 and x.ndim == 2
    ss = np.einsum("ij,ij->i", x, x, dtype=np.float32)
    nz = ss > 0
    inv = np.ones_like(ss)
    inv[nz] = np.float32(1.0) / np.sqrt(ss[nz], dtype=np.float32)
    x *= inv[:, None]
    return x


def topk_desc(scores: np.ndarray, k: int, extra: int = 0):
    """Row-wise top-(k+extra) of a [Q,N] fp32 matrix in canonical order (-score, label).
    Returns (D [Q,k+extra], I [Q,k+extra]); slots past N hold (-FLT_MAX, -1)."""
    Q, N = scores.shape
    kk = k + extra
    D = np.full((Q, kk), NEG_FLT_MAX, dtype=np.float32)
    I = np.full((Q, kk), -1, dtype=np.int64)
    take = min(kk, N)
    if take == 0:
        return D, I
    for qi in range(Q):
        row = scores[qi]
        if take < N:
            part = np.argpartition(-row, take - 1)[:take]
            kth = row[part].min()
            cand = np.nonzero(row >= kth)[0]  # includes every tie of the boundary value
        else:
            cand = np.arange(N)
        order = np.lexsort((cand, -row[cand]))[:take]
        sel = cand[order]
        D[qi, :take] = row[sel]
        I[qi, :take] = sel
    return D, I


class OracleIndexFlatIP:
    """faiss.IndexFlatIP stand-in: exact fp32 inner products."""

    is_trained = True

    def __init__(self, d: int):
        self.d = d
        self.xb = np.zeros((0, d), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return self.xb.shape[0]

    def train(self, x):  # no-op, like faiss
        pass

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.shape[1] == self.d
        self.xb = np.concatenate([self.xb, x], axis=0)

    def scores(self, q: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(q, dtype=np.float32) @ self.xb.T

    def search(self, q: np.ndarray, k: int, extra: int = 0):
        return topk_desc(self.scores(q), k, extra)


class OracleIndexFlatL2(OracleIndexFlatIP):
    """Exact squared-L2 search (what `faiss.IndexHNSWFlat` approximates): direct fp32 sum of squared
    differences, top-k smallest in canonical (distance, label) order."""

    def distances(self, q: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.empty((q.shape[0], self.ntotal), dtype=np.float32)
        for i in range(q.shape[0]):
            diff = self.xb - q[i]
            out[i] = np.einsum("ij,ij->i", diff, diff, dtype=np.float32)
        return out

    def search(self, q: np.ndarray, k: int, extra: int = 0):
        D, I = topk_desc(-self.distances(q), k, extra)
        D = -D
        D[I < 0] = -NEG_FLT_MAX
        return D, I


class OracleFAISSIndex:
    """The reference's `FAISSIndex` wrapper (faiss_retrieval.py:14-256) over oracle indexes."""

    def __init__(self, dimension, index_type='IVF', nlist=100, nprobe=10, use_gpu=False, **kw):
        self.dimension = dimension
        self.index_type = index_type
        self.nlist = nlist
        self.nprobe = nprobe
        self.use_gpu = use_gpu
        self.id_map = []
        if index_type == 'Flat':
            self.index = OracleIndexFlatIP(dimension)
        elif index_type == 'IVF':
            from .ivf import OracleIndexIVFFlat
            self.index = OracleIndexIVFFlat(dimension, nlist, **kw)
        elif index_type == 'IVFPQ':
            from .ivfpq import OracleIndexIVFPQ
            self.index = OracleIndexIVFPQ(dimension, nlist, kw.pop('pq_m', 8), 8, **kw)
        elif index_type == 'HNSW':
            self.index = OracleIndexFlatL2(dimension)
        else:
            raise ValueError(f"Unknown index type: {self.index_type}")

    def train(self, embeddings):
        if not self.index.is_trained:
            self.index.train(embeddings.astype('float32'))

    def add(self, embeddings, ad_ids=None):
        if not self.index.is_trained:
            self.train(embeddings)                      # on the raw input (faiss_retrieval.py:107-108)
        embeddings = embeddings.astype('float32')       # fresh copy (:114)
        normalize_L2(embeddings)                        # (:115)
        self.index.add(embeddings)                      # (:118)
        if ad_ids is None:
            ad_ids = list(range(len(self.id_map), len(self.id_map) + len(embeddings)))
        self.id_map.extend(ad_ids)

    def search(self, query_embeddings, k=100, return_distances=True, extra=0):
        q = query_embeddings.astype('float32')
        normalize_L2(q)
        if hasattr(self.index, 'nprobe'):
            self.index.nprobe = self.nprobe
        distances, indices = self.index.search(q, k, extra) if extra else self.index.search(q, k)
        indices = np.asarray(indices)
        table = np.empty(len(self.id_map), dtype=object)
        table[:] = self.id_map
        ad_ids = np.array(table[indices].tolist())      # id_map[idx]; -1 wraps to the last id (:159-160)
        if return_distances:
            return ad_ids, distances
        return ad_ids

    def batch_search(self, query_embeddings, k=100, batch_size=1000):
        ids, ds = [], []
        for i in range(0, len(query_embeddings), batch_size):
            a, d = self.search(query_embeddings[i:i + batch_size], k)
            ids.append(a)
            ds.append(d)
        return np.vstack(ids), np.vstack(ds)

    def get_stats(self):
        return {'index_type': self.index_type, 'dimension': self.dimension,
                'num_vectors': self.index.ntotal, 'is_trained': self.index.is_trained,
                'nlist': self.nlist, 'nprobe': self.nprobe}
