"""IVF index families on the B200 kernels (faiss-shaped objects used by `FAISSIndex.index`).

`IndexIVFFlat` mirrors `faiss.IndexIVFFlat(IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT)`
(faiss_retrieval.py:50-55): `is_trained`, `ntotal`, `nprobe`, `train`, `add`, `search`.
Extra (parity plumbing): `export_centroids` / `import_centroids` share the coarse
quantiser with the oracle, `list_sizes` exposes the inverted-list histogram.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .faiss_retrieval import METRIC_INNER_PRODUCT, METRIC_L2, _DeviceIndex


class IndexIVFFlat(_DeviceIndex):
    kind = _lib.KIND_IVF_FLAT
    metric = METRIC_INNER_PRODUCT
    _supports_retry = False

    def __init__(self, d: int, nlist: int, *, device=None, pq_m: int = 0):
        self.nlist = int(nlist)
        self.nprobe = 1  # faiss default; FAISSIndex overwrites it before every search (faiss_retrieval.py:150-151)
        super().__init__(d, nlist=nlist, pq_m=pq_m, pq_bits=8 if pq_m else 0, device=device)

    def search(self, x, k: int, *, normalize: bool = False, nprobe: int = 0, return_device: bool = False):
        return super().search(x, k, normalize=normalize, nprobe=nprobe or self.nprobe, return_device=return_device)

    def search_device(self, q, k: int, *, normalize: bool = False, nprobe: int = 0, tau=None, want_status: bool = True):
        return super().search_device(q, k, normalize=normalize, nprobe=nprobe or self.nprobe, tau=None,
                                     want_status=want_status)

    # ---- parity plumbing -----------------------------------------------------------
    def export_centroids(self) -> np.ndarray:
        out = np.empty((self.nlist, self.d), dtype=np.float32)
        _lib.check(self._lib.b2r_index_export_centroids(self._h, out.ctypes.data))
        return out

    def import_centroids(self, centroids) -> None:
        c = np.ascontiguousarray(centroids, dtype=np.float32)
        if c.shape != (self.nlist, self.d):
            raise ValueError(f"centroids must be [{self.nlist}, {self.d}]")
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b2r_index_import_centroids(self._h, c.ctypes.data))

    def list_sizes(self) -> np.ndarray:
        out = np.zeros(self.nlist, dtype=np.int64)
        _lib.check(self._lib.b2r_index_list_sizes(self._h, out.ctypes.data))
        return out

    def state_dict(self) -> dict:
        return {"centroids": self.export_centroids()} if self.is_trained else {}

    def load_state_dict(self, state: dict) -> None:
        if "centroids" in state:
            self.import_centroids(state["centroids"])


class IndexIVFPQ(IndexIVFFlat):
    """faiss.IndexIVFPQ(IndexFlatIP(d), d, nlist, m, 8) — default metric L2 (faiss_retrieval.py:57-63)."""
    kind = _lib.KIND_IVF_PQ
    metric = METRIC_L2

    def __init__(self, d: int, nlist: int, m: int = 8, *, device=None):
        self.pq_m = int(m)
        super().__init__(d, nlist, device=device, pq_m=m)


def create(owner, kind):
    if kind == 'IVF':
        return IndexIVFFlat(owner.dimension, owner.nlist, device=owner._device)
    if kind == 'IVFPQ':
        return IndexIVFPQ(owner.dimension, owner.nlist, owner._pq_m, device=owner._device)
    raise ValueError(f"Unknown index type: {kind}")
