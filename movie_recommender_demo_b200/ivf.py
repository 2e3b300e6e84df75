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
    # The fused list scan (sample pass -> per-query threshold -> filter in the scan's epilogue) can leave a query
    # flagged (too few candidates / candidate buffer overflow / threshold above the provable rescore window).
    # Flagged queries are re-run through the dump path, whose exact in-kernel fallback always succeeds.
    _supports_retry = True
    _collective_retry = False    # sharded.py: the caller-threshold retry protocol is the flat index's only

    def _retry(self, qt, k, normalize, nprobe, st, tau_retry, D_dev, I_dev, D_np, I_np):
        from .faiss_retrieval import _RETRY_BITS
        torch = self._torch
        bad = np.nonzero(st & _RETRY_BITS)[0]
        self.last_retries = 0
        if bad.size == 0 or int(self.get_param("ivf_fused")) == 0:
            return st
        bad_t = torch.as_tensor(bad, device=self.device)
        self.set_param("ivf_fused", 0)
        try:
            D2, I2, st2, _ = self._search_prepared(qt.index_select(0, bad_t), k, normalize=normalize, nprobe=nprobe)
        finally:
            self.set_param("ivf_fused", 1)
        if D_dev is not None:
            D_dev.index_copy_(0, bad_t, D2)
            I_dev.index_copy_(0, bad_t, I2)
        else:
            D_np[bad] = D2.cpu().numpy()
            I_np[bad] = I2.cpu().numpy()
        st[bad] = st2.cpu().numpy()
        self.last_retries = 1
        return st

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
        out = np.empty((self.nlist, self._dp), dtype=np.float32)
        _lib.check(self._lib.b2r_index_export_centroids(self._h, out.ctypes.data))
        return out if self._dp == self.d else np.ascontiguousarray(out[:, :self.d])

    def import_centroids(self, centroids) -> None:
        self._graphs = {}
        c = np.ascontiguousarray(centroids, dtype=np.float32)
        if c.shape != (self.nlist, self.d):
            raise ValueError(f"centroids must be [{self.nlist}, {self.d}]")
        if self._dp != self.d:
            c = np.ascontiguousarray(np.pad(c, ((0, 0), (0, self._dp - self.d))))
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b2r_index_import_centroids(self._h, c.ctypes.data))

    def list_sizes(self) -> np.ndarray:
        out = np.zeros(self.nlist, dtype=np.int64)
        _lib.check(self._lib.b2r_index_list_sizes(self._h, out.ctypes.data))
        return out

    def lists_by_label(self) -> np.ndarray:
        """int64 [ntotal] inverted-list id of every vector, in insertion (label) order."""
        torch = self._torch
        n = self.ntotal
        stored = np.repeat(np.arange(self.nlist, dtype=np.int64), self.list_sizes())   # rows are sorted by list
        with torch.cuda.device(self.device):
            labels = torch.empty(n, dtype=torch.int64, device=self.device)
            _lib.check(self._lib.b2r_index_get_labels(self._h, 0, n, labels.data_ptr(),
                                                      int(torch.cuda.current_stream(self.device).cuda_stream)))
        out = np.empty(n, dtype=np.int64)
        out[labels.cpu().numpy()] = stored
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

    _supports_retry = False     # ADC scan: no fused filter, the threshold kernel's exact fallback is in-kernel

    def __init__(self, d: int, nlist: int, m: int = 8, *, device=None):
        self.pq_m = int(m)
        super().__init__(d, nlist, device=device, pq_m=m)

    def export_codebooks(self) -> np.ndarray:
        out = np.empty((self.pq_m, 256, self.d // self.pq_m), dtype=np.float32)
        _lib.check(self._lib.b2r_index_export_codebooks(self._h, out.ctypes.data))
        return out

    def import_codebooks(self, codebooks) -> None:
        self._graphs = {}
        cb = np.ascontiguousarray(codebooks, dtype=np.float32)
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b2r_index_import_codebooks(self._h, cb.ctypes.data))

    def codes_by_label(self) -> np.ndarray:
        """uint8 [ntotal, m] PQ codes in insertion (label) order."""
        torch = self._torch
        n = self.ntotal
        with torch.cuda.device(self.device):
            sp = int(torch.cuda.current_stream(self.device).cuda_stream)
            codes = torch.empty((n, self.pq_m), dtype=torch.uint8, device=self.device)
            labels = torch.empty(n, dtype=torch.int64, device=self.device)
            _lib.check(self._lib.b2r_index_get_codes(self._h, 0, n, codes.data_ptr(), sp))
            _lib.check(self._lib.b2r_index_get_labels(self._h, 0, n, labels.data_ptr(), sp))
            out = torch.empty_like(codes)
            out[labels] = codes
        return out.cpu().numpy()

    def reconstruct_n(self, i0: int, n: int):
        raise NotImplementedError("IVF-PQ stores codes only")

    def add_codes(self, codes, lists) -> None:
        self._graphs = {}
        torch = self._torch
        c = torch.as_tensor(np.ascontiguousarray(codes, dtype=np.uint8)).to(self.device)
        l = torch.as_tensor(np.ascontiguousarray(lists, dtype=np.int64)).to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b2r_index_add_codes(self._h, c.shape[0], c.data_ptr(), l.data_ptr(),
                                                     int(torch.cuda.current_stream(self.device).cuda_stream)))
            torch.cuda.current_stream(self.device).synchronize()

    stores_vectors = False

    def state_dict(self) -> dict:
        if not self.is_trained:
            return {}
        st = {"centroids": self.export_centroids(), "codebooks": self.export_codebooks()}
        if self.ntotal:
            st["codes"] = self.codes_by_label()
            st["lists"] = self.lists_by_label()
        return st

    def load_state_dict(self, state: dict) -> None:
        if "codebooks" in state:
            self.import_codebooks(state["codebooks"])
        if "centroids" in state:
            self.import_centroids(state["centroids"])
        if "codes" in state:
            self.add_codes(state["codes"], state["lists"])


def create(owner, kind):
    if kind == 'IVF':
        index = IndexIVFFlat(owner.dimension, owner.nlist, device=owner._device)
    elif kind == 'IVFPQ':
        index = IndexIVFPQ(owner.dimension, owner.nlist, owner._pq_m, device=owner._device)
    else:
        raise ValueError(f"Unknown index type: {kind}")
    index.nprobe = owner.nprobe   # the wrapper re-applies it before every search anyway (faiss_retrieval.py:150-151)
    return index
