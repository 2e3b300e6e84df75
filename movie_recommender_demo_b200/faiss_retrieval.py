"""Drop-in for the reference `faiss_retrieval` module, backed by libb2retr.so on a B200.

Same public surface as the reference (faiss_retrieval.py:14-369): `FAISSIndex` with
`train / add / search / batch_search / save / load / get_stats` and the attributes
`dimension, index_type, nlist, nprobe, use_gpu, index, id_map`; `TwoStageRetriever`.
`FAISSIndex.index` is a faiss-shaped object (`ntotal`, `is_trained`, `nprobe`, `train`,
`add`, `search -> (D, I)`), because callers read `faiss_index.index.ntotal`
(train.py:231, inference.py:156).

What changes underneath: vectors live in HBM (fp32 master + bf16 scan copy), normalisation,
the score contraction (tcgen05), top-k selection, fp32 re-scoring and the id remap
(`id_map[idx]`, faiss_retrieval.py:159-160) all run in hand-written sm_100a kernels.
There is no CPU path: without a CUDA device / the built library every call raises.

Documented deviations from the reference:
  * `use_gpu` is accepted and ignored - the index is always on the GPU.
  * 'HNSW' (faiss_retrieval.py:65-70) is served by an EXACT squared-L2 scan on the flat tcgen05 path
    (`IndexHNSWFlat` below): same surface and ordering (ascending L2), recall 1.0 instead of the graph's
    approximation - on a B200 the brute-force scan of 1M x 256 rows (~0.13 ms) is faster than CPU graph
    traversal, so no graph is built.
  * extra keyword-only constructor arguments with reference-preserving defaults:
    `device`, `pq_m`.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import struct
import time
import warnings
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib

__all__ = ["benchmark_faiss_index", "FAISSIndex", "TwoStageRetriever", "IndexFlatIP", "IndexHNSWFlat",
           "METRIC_INNER_PRODUCT", "METRIC_L2"]

METRIC_INNER_PRODUCT = _lib.METRIC_IP
METRIC_L2 = _lib.METRIC_L2
_RETRY_BITS = _lib.ST_TOO_FEW | _lib.ST_NEED_LOWER_TAU | _lib.ST_CAND_OVERFLOW
_MAX_RETRIES = 8
_PIPE_CHUNK = 1024   # queries per chunk when device->host result copies are pipelined
_GRAPH_MAX_Q = 256   # host-result searches up to this batch replay a captured CUDA graph (launch-bound regime)
_GRAPH_CACHE = 16    # captured (batch, k, nprobe) shapes kept per index


def _pipe_chunks(nq: int):
    """Chunk boundaries for the pipelined host-result path.  Big chunks scan faster (more queries share
    one pass over the corpus), but the LAST chunk's device->host copy cannot hide behind anything, so the
    sizes shrink towards the tail: the last chunk is _PIPE_CHUNK queries and each earlier one at most 4x
    its successor (a chunk's result copy takes ~1/5 of the same chunk's search, so it hides under the next)."""
    import os
    forced = os.environ.get("B2R_PIPE_SIZES")            # measurement hook: "3584,512" = explicit chunk sizes
    if forced:
        sizes = [int(v) for v in forced.split(",")]
        if sum(sizes) == nq and all(v > 0 for v in sizes):
            out, lo = [], 0
            for v in sizes:
                out.append((lo, lo + v))
                lo += v
            return out
    sizes = [min(nq, _PIPE_CHUNK)]
    left = nq - sizes[0]
    while left > 0:
        take = min(left, 4 * sizes[0])
        sizes.insert(0, take)
        left -= take
    out, lo = [], 0
    for s in sizes:
        out.append((lo, lo + s))
        lo += s
    return out


def _stream_ptr(torch, device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class _DeviceIndex:
    """faiss-shaped handle over a `b2r_index` (one GPU)."""

    kind = _lib.KIND_FLAT
    metric = METRIC_INNER_PRODUCT
    _supports_retry = True   # caller-threshold retries exist for the flat scan only

    def __init__(self, d: int, *, nlist: int = 0, pq_m: int = 0, pq_bits: int = 0, device=None):
        torch = _lib.require_cuda()
        self._torch = torch
        self._lib = _lib.load()
        self.d = int(d)
        # The kernels tile the contraction dimension in 64-element (128-byte, swizzled) chunks.  Other
        # dimensions are zero-padded to the next multiple of 64 on the way in (inner products, norms and
        # centroids are unchanged by zero columns) and sliced back on the way out.
        self._dp = (self.d + 63) // 64 * 64
        if not (1 <= self.d <= 256):
            raise ValueError(f"dimension {d} not supported: the B200 scan kernels hold a 128-query block of up to "
                             "256 dimensions in shared memory")
        if pq_m and self._dp != self.d:
            raise ValueError("IVFPQ needs a dimension that is a multiple of 64 (sub-quantiser slices must not "
                             "straddle padding)")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.init()
        torch.zeros(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        _lib.check(self._lib.b2r_index_create(C.byref(h), self.kind, self._dp, int(nlist), int(pq_m),
                                              int(pq_bits), self.metric, self.device.index))
        self._h = h
        self._ws = None
        self._ids_set = False
        self.last_status = None  # per-query status bits of the most recent search (numpy)
        self.last_retries = 0
        self._copy_stream = None
        self._graphs = {}        # (nq, k, normalize, nprobe) -> captured search (small batches), see _graph_entry
        self.replayed_launches = 0   # kernels launched through graph replays (they bypass the library's own counter)

    # ------------------------------------------------------------ lifetime
    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.b2r_index_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---------------------------------------------------------- properties
    @property
    def ntotal(self) -> int:
        return int(self._lib.b2r_index_ntotal(self._h))

    @property
    def is_trained(self) -> bool:
        return bool(self._lib.b2r_index_is_trained(self._h))

    def set_param(self, name: str, value: float) -> None:
        self._graphs = {}
        _lib.check(self._lib.b2r_index_set_param(self._h, name.encode(), float(value)))

    def get_param(self, name: str) -> float:
        return float(self._lib.b2r_index_get_param(self._h, name.encode()))

    # ------------------------------------------------------------- helpers
    def _to_device_f32(self, x, what: str):
        """numpy (any float) or torch tensor -> contiguous fp32 CUDA tensor [n, d] (a copy
        only when needed; the caller's array is never written)."""
        torch = self._torch
        if isinstance(x, torch.Tensor):
            t = x.detach()
            if t.dim() != 2 or t.shape[1] != self.d:
                raise ValueError(f"{what}: expected shape [n, {self.d}], got {tuple(t.shape)}")
            return self._pad(t.to(device=self.device, dtype=torch.float32).contiguous())
        a = np.asarray(x)
        if a.ndim != 2 or a.shape[1] != self.d:
            raise ValueError(f"{what}: expected shape [n, {self.d}], got {a.shape}")
        a = np.ascontiguousarray(a, dtype=np.float32)
        if not a.flags.writeable:       # read-only views (np.frombuffer, memmap): torch wants a writable source
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", UserWarning)
                return self._pad(torch.from_numpy(a).to(self.device, non_blocking=False))
        return self._pad(torch.from_numpy(a).to(self.device, non_blocking=False))

    def _pad(self, t):
        if self._dp == self.d:
            return t
        return self._torch.nn.functional.pad(t, (0, self._dp - self.d))

    def _workspace(self, nbytes: int):
        torch = self._torch
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=self.device)
        return self._ws

    # ---------------------------------------------------------- faiss API
    def train(self, x) -> None:
        self._graphs = {}
        xt = self._to_device_f32(x, "train")
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b2r_index_train(self._h, xt.shape[0], xt.data_ptr(), 1234,
                                                 _stream_ptr(self._torch, self.device)))

    def add(self, x, normalize: bool = False) -> None:
        """faiss `index.add(x)`; `normalize=True` fuses faiss.normalize_L2 into the ingest."""
        self._graphs = {}
        xt = self._to_device_f32(x, "add")
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b2r_index_add(self._h, xt.shape[0], xt.data_ptr(), int(bool(normalize)),
                                               _stream_ptr(self._torch, self.device)))
        self._ids_set = False

    def reserve(self, rows: int) -> None:
        """Pre-size the corpus buffers (flat index): adds up to `rows` vectors then never reallocate."""
        self._graphs = {}
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b2r_index_reserve(self._h, int(rows), _stream_ptr(self._torch, self.device)))

    def reset(self) -> None:
        self._graphs = {}
        _lib.check(self._lib.b2r_index_reset(self._h))
        self._ids_set = False

    def set_ids(self, ids) -> None:
        """Device-side id map: search returns ids[label] (None clears)."""
        self._graphs = {}
        torch = self._torch
        with torch.cuda.device(self.device):
            sp = _stream_ptr(torch, self.device)
            if ids is None:
                _lib.check(self._lib.b2r_index_set_ids(self._h, 0, None, sp))
                self._ids_set = False
                return
            t = ids if isinstance(ids, torch.Tensor) else torch.as_tensor(np.asarray(ids, dtype=np.int64))
            t = t.to(device=self.device, dtype=torch.int64).contiguous()
            _lib.check(self._lib.b2r_index_set_ids(self._h, t.numel(), t.data_ptr(), sp))
            torch.cuda.current_stream(self.device).synchronize()
            self._ids_set = True

    def set_label_base(self, base: int) -> None:
        self._graphs = {}
        _lib.check(self._lib.b2r_index_set_label_base(self._h, int(base)))

    def search_device(self, q, k: int, *, normalize: bool = False, nprobe: int = 0, tau=None,
                      want_status: bool = True):
        """Asynchronous search on device tensors. Returns (D, I, status, tau_retry) CUDA tensors."""
        return self._search_prepared(self._to_device_f32(q, "search"), k, normalize=normalize, nprobe=nprobe,
                                     tau=tau, want_status=want_status)

    def _search_prepared(self, qt, k: int, *, normalize: bool = False, nprobe: int = 0, tau=None,
                         want_status: bool = True):
        """`search_device` on queries that already went through `_to_device_f32` (fp32, on device, padded)."""
        torch = self._torch
        nq = qt.shape[0]
        k = int(k)
        with torch.cuda.device(self.device):
            D = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            status = torch.zeros(nq, dtype=torch.int32, device=self.device) if want_status else None
            tau_retry = torch.empty(nq, dtype=torch.float32, device=self.device) if want_status else None
            if nq == 0:
                return D, I, status, tau_retry
            need = int(self._lib.b2r_index_search_workspace(self._h, nq, k, int(nprobe)))
            ws = self._workspace(need)
            _lib.check(self._lib.b2r_index_search(
                self._h, nq, qt.data_ptr(), int(bool(normalize)), k, int(nprobe), D.data_ptr(), I.data_ptr(),
                status.data_ptr() if status is not None else None,
                tau_retry.data_ptr() if tau_retry is not None else None,
                tau.data_ptr() if tau is not None else None,
                ws.data_ptr(), ws.numel(), _stream_ptr(torch, self.device)))
        return D, I, status, tau_retry

    def search(self, x, k: int, *, normalize: bool = False, nprobe: int = 0, return_device: bool = False):
        """faiss `index.search(x, k) -> (D, I)`; numpy in, numpy out (CUDA tensors accepted).

        Queries the kernels flag as "not provably exact" (threshold too high / candidate
        overflow) are re-run with the threshold the device suggests; that needs the status on
        the host, which rides along with the result copy."""
        torch = self._torch
        self.last_retries = 0
        nq = len(x)
        ent = None
        if not return_device and 0 < nq <= _GRAPH_MAX_Q and not os.environ.get("B2R_NO_GRAPHS"):
            ent = self._graph_entry(nq, int(k), bool(normalize), int(nprobe))
        if ent:
            return self._search_graph(ent, x, k, normalize, nprobe)
        qt = self._to_device_f32(x, "search")
        if not return_device and qt.shape[0] >= 2 * _PIPE_CHUNK:
            return self._search_pipelined(qt, k, normalize, nprobe)
        D, I, status, tau_retry = self._search_prepared(qt, k, normalize=normalize, nprobe=nprobe)
        # results + status ride to pinned host buffers in one batch of async copies, one sync
        st_h = self._to_pinned(status)
        D_h = I_h = None
        if not return_device:
            D_h, I_h = self._to_pinned(D), self._to_pinned(I)
        torch.cuda.current_stream(self.device).synchronize()
        st = st_h.numpy()
        if (st & _RETRY_BITS).any() and self._supports_retry:
            st = self._retry(qt, k, normalize, nprobe, st.copy(), tau_retry, D, I, None, None)
            D_h = None  # stale
        self.last_status = st
        self._warn_status(st)
        if return_device:
            return D, I
        if D_h is None:
            D_h, I_h = self._to_pinned(D), self._to_pinned(I)
            torch.cuda.current_stream(self.device).synchronize()
        return D_h.numpy(), I_h.numpy()

    # ---- small batches: the launch sequence of one search is fixed for a given (batch, k, nprobe) and a
    #      given corpus state, and at batch <= 256 the gaps between its 6-14 short launches are a third of
    #      the step -> capture it once in a CUDA graph (static query / result buffers), replay afterwards.
    def _graph_entry(self, nq: int, k: int, normalize: bool, nprobe: int):
        key = (nq, k, normalize, nprobe)
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        if self.ntotal == 0:
            return None
        torch = self._torch
        with torch.cuda.device(self.device):
            need = int(self._lib.b2r_index_search_workspace(self._h, nq, k, nprobe))
            # I | D | status live in ONE buffer (labels first: 8-byte aligned) so that the result leaves for the
            # host in one copy instead of three
            nI, nD = nq * k * 8, nq * k * 4
            packed = torch.zeros(nI + nD + nq * 4, dtype=torch.uint8, device=self.device)
            ent = {"q": torch.zeros((nq, self._dp), dtype=torch.float32, device=self.device),
                   "packed": packed, "cuts": (nI, nI + nD),
                   "I": packed[:nI].view(torch.int64).view(nq, k),
                   "D": packed[nI:nI + nD].view(torch.float32).view(nq, k),
                   "status": packed[nI + nD:].view(torch.int32),
                   "tau_retry": torch.empty(nq, dtype=torch.float32, device=self.device),
                   "ws": torch.empty(max(need, 1), dtype=torch.uint8, device=self.device)}

            def launch():
                _lib.check(self._lib.b2r_index_search(
                    self._h, nq, ent["q"].data_ptr(), int(normalize), k, nprobe, ent["D"].data_ptr(),
                    ent["I"].data_ptr(), ent["status"].data_ptr(), ent["tau_retry"].data_ptr(), None,
                    ent["ws"].data_ptr(), ent["ws"].numel(), _stream_ptr(torch, self.device)))

            n0 = int(self._lib.b2r_debug_launch_count())
            launch()    # eager once: one-time kernel attribute set-up must not happen under capture
            ent["launches"] = int(self._lib.b2r_debug_launch_count()) - n0
            torch.cuda.current_stream(self.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    launch()
                ent["graph"] = graph
            except Exception:      # not capturable on this build/driver: stay on the eager path for this shape
                ent = False
                torch.cuda.synchronize(self.device)
        if len(self._graphs) >= _GRAPH_CACHE:
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = ent
        return ent

    def search_device_static(self, q, k: int, *, normalize: bool = False, nprobe: int = 0):
        """Device-resident small-batch search through the captured graph: (D, I, status) are the graph's
        STATIC output tensors, valid until the next search on this index.  Falls back to `search_device`
        when the shape is not graph-eligible.  No retry handling: the caller reads `status`."""
        torch = self._torch
        nq = len(q)
        ent = self._graph_entry(nq, int(k), bool(normalize), int(nprobe)) if 0 < nq <= _GRAPH_MAX_Q else None
        if not ent:
            D, I, st, _ = self.search_device(q, k, normalize=normalize, nprobe=nprobe)
            return D, I, st
        src = q.detach() if isinstance(q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
        if src.dim() != 2 or src.shape[1] != self.d:
            raise ValueError(f"search: expected shape [n, {self.d}], got {tuple(src.shape)}")
        with torch.cuda.device(self.device):
            buf = ent["q"]
            (buf if self._dp == self.d else buf[:, :self.d]).copy_(src, non_blocking=True)
            ent["graph"].replay()
            self.replayed_launches += ent["launches"]
        return ent["D"], ent["I"], ent["status"]

    def _search_graph(self, ent, x, k, normalize, nprobe):
        torch = self._torch
        q = ent["q"]
        if isinstance(x, torch.Tensor):
            src = x.detach()
            if src.dim() != 2 or src.shape[1] != self.d:
                raise ValueError(f"search: expected shape [n, {self.d}], got {tuple(src.shape)}")
        else:
            a = np.asarray(x)
            if a.ndim != 2 or a.shape[1] != self.d:
                raise ValueError(f"search: expected shape [n, {self.d}], got {a.shape}")
            a = np.ascontiguousarray(a, dtype=np.float32)
            if not a.flags.writeable:
                a = a.copy()
            src = torch.from_numpy(a)
        with torch.cuda.device(self.device):
            (q if self._dp == self.d else q[:, :self.d]).copy_(src, non_blocking=True)
            ent["graph"].replay()
            self.replayed_launches += ent["launches"]
            D, I = ent["D"], ent["I"]
            nq = q.shape[0]
            cI, cD = ent["cuts"]
            host = self._to_pinned(ent["packed"])      # labels | distances | status in one async copy
            torch.cuda.current_stream(self.device).synchronize()
            st = host[cD:].view(torch.int32).numpy()
            if (st & _RETRY_BITS).any() and self._supports_retry:
                st = self._retry(q, k, normalize, nprobe, st.copy(), ent["tau_retry"], D, I, None, None)
                host = self._to_pinned(ent["packed"])
                torch.cuda.current_stream(self.device).synchronize()
        self.last_status = st
        self._warn_status(st)
        return (host[cI:cD].view(torch.float32).view(nq, int(k)).numpy(),
                host[:cI].view(torch.int64).view(nq, int(k)).numpy())

    def _search_pipelined(self, qt, k, normalize, nprobe):
        """Large batches: search in chunks of _PIPE_CHUNK queries and copy each chunk's results to
        pinned host memory on a side stream while the next chunk is being searched."""
        torch = self._torch
        nq = qt.shape[0]
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        side = self._copy_stream
        D_h = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
        I_h = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
        st_h = torch.empty(nq, dtype=torch.int32, pin_memory=True)
        keep, taus = [], []
        for lo, hi in _pipe_chunks(nq):
            D, I, st, tr = self._search_prepared(qt[lo:hi], k, normalize=normalize, nprobe=nprobe)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                D_h[lo:hi].copy_(D, non_blocking=True)
                I_h[lo:hi].copy_(I, non_blocking=True)
                st_h[lo:hi].copy_(st, non_blocking=True)
            keep.append((D, I, st))     # alive until the side stream is drained
            taus.append(tr)
        side.synchronize()
        st = st_h.numpy()
        D_np, I_np = D_h.numpy(), I_h.numpy()
        if (st & _RETRY_BITS).any() and self._supports_retry:
            st = self._retry(qt, k, normalize, nprobe, st.copy(), torch.cat(taus), None, None, D_np, I_np)
        self.last_status = st
        self._warn_status(st)
        return D_np, I_np

    def _retry(self, qt, k, normalize, nprobe, st, tau_retry, D_dev, I_dev, D_np, I_np):
        """Re-run the queries the kernels flagged as not provably exact, with the thresholds the
        device suggested, until none is flagged or no progress is made."""
        torch = self._torch
        retries = 0
        prev_tau = None
        while retries < _MAX_RETRIES:
            bad = np.nonzero(st & _RETRY_BITS)[0]
            if bad.size == 0:
                break
            bad_t = torch.as_tensor(bad, device=self.device)
            tau = tau_retry.index_select(0, bad_t).contiguous()
            tau_host = tau.cpu().numpy()
            if prev_tau is not None and prev_tau.shape == tau_host.shape and np.array_equal(prev_tau, tau_host):
                break  # no progress (e.g. a tie group larger than the candidate buffer)
            prev_tau = tau_host
            D2, I2, st2, tr2 = self._search_prepared(qt.index_select(0, bad_t), k, normalize=normalize,
                                                  nprobe=nprobe, tau=tau)
            if D_dev is not None:
                D_dev.index_copy_(0, bad_t, D2)
                I_dev.index_copy_(0, bad_t, I2)
            else:
                D_np[bad] = D2.cpu().numpy()
                I_np[bad] = I2.cpu().numpy()
            tau_retry.index_copy_(0, bad_t, tr2)
            st[bad] = st2.cpu().numpy()
            retries += 1
        self.last_retries = retries
        return st

    def _warn_status(self, st) -> None:
        if (st != 0).any():
            warnings.warn(f"b200 search: {int((st != 0).sum())} queries not provably exact "
                          f"(status bits {sorted(set(int(s) for s in st if s))})")

    def _to_pinned(self, t):
        """async device->pinned-host copy (torch's caching host allocator recycles the blocks;
        the returned numpy view keeps its block alive)."""
        h = self._torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        return h

    def reconstruct_n(self, i0: int, n: int):
        """fp32 master rows of labels [i0, i0+n) as a CUDA tensor (faiss `reconstruct_n`)."""
        torch = self._torch
        with torch.cuda.device(self.device):
            sp = _stream_ptr(torch, self.device)
            if self.kind == _lib.KIND_FLAT:
                out = torch.empty((n, self._dp), dtype=torch.float32, device=self.device)
                _lib.check(self._lib.b2r_index_get_vectors(self._h, int(i0), int(n), out.data_ptr(), sp))
                return out if self._dp == self.d else out[:, :self.d].contiguous()
            # IVF stores rows sorted by list: fetch everything, undo the permutation
            nt = self.ntotal
            rows = torch.empty((nt, self._dp), dtype=torch.float32, device=self.device)
            labels = torch.empty(nt, dtype=torch.int64, device=self.device)
            _lib.check(self._lib.b2r_index_get_vectors(self._h, 0, nt, rows.data_ptr(), sp))
            _lib.check(self._lib.b2r_index_get_labels(self._h, 0, nt, labels.data_ptr(), sp))
            out = torch.empty_like(rows)
            out[labels] = rows
            return out[i0:i0 + n, :self.d].contiguous()

    # test-only: the full bf16 score matrix via the tcgen05 dump mode / a CUDA-core loop
    def debug_scores(self, x, impl: str = "tc", normalize: bool = False):
        torch = self._torch
        qt = self._to_device_f32(x, "debug_scores")
        out = torch.empty((qt.shape[0], self.ntotal), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            sp = _stream_ptr(torch, self.device)
            if impl == "tc":
                ws = self._workspace(((qt.shape[0] + 255) // 256 * 256) * self._dp * 8 + (1 << 16))
                _lib.check(self._lib.b2r_debug_scores_tc(self._h, qt.shape[0], qt.data_ptr(), int(normalize),
                                                         out.data_ptr(), ws.data_ptr(), ws.numel(), sp))
            else:
                _lib.check(self._lib.b2r_debug_scores_simt(self._h, qt.shape[0], qt.data_ptr(), int(normalize),
                                                           out.data_ptr(), sp))
        return out


class IndexFlatIP(_DeviceIndex):
    """Exact inner-product index (faiss.IndexFlatIP, faiss_retrieval.py:48)."""
    kind = _lib.KIND_FLAT
    metric = METRIC_INNER_PRODUCT


class _HNSWParams:
    """`index.hnsw` of faiss.IndexHNSWFlat: the reference sets efConstruction / efSearch on it
    (faiss_retrieval.py:69-70).  Kept as plain attributes; an exact scan has no use for them."""

    def __init__(self, M: int):
        self.M = int(M)
        self.efConstruction = 40
        self.efSearch = 16


class IndexHNSWFlat(_DeviceIndex):
    """Stand-in for `faiss.IndexHNSWFlat(d, M)` (faiss_retrieval.py:65-70; faiss default metric L2,
    distances ascending): an EXACT squared-L2 search.  HNSW approximates the L2 nearest neighbours; this
    returns them exactly (what HNSW returns at recall 1.0), from the same kernels as `IndexFlatIP`.

    How: every stored row has unit norm (the reference wrapper normalises on add, faiss_retrieval.py:115;
    a raw `add` of other vectors is rejected), so ||q - x||^2 = ||q||^2 + 1 - 2<q, x> is monotone in the
    inner product: the flat IP top-k (fp32-rescored, canonical tie order) IS the L2 top-k, and the distances
    follow from the exact fp32 inner products.  Missing slots: label -1, distance +FLT_MAX as in faiss."""
    kind = _lib.KIND_FLAT
    metric = METRIC_INNER_PRODUCT     # metric of the underlying scan; results are reported as L2
    reported_metric = METRIC_L2
    _NORM_TOL = 1e-3

    def __init__(self, d: int, M: int = 32, *, device=None):
        super().__init__(d, device=device)
        self.hnsw = _HNSWParams(M)

    def add(self, x, normalize: bool = False) -> None:
        xt = self._to_device_f32(x, "add")
        if xt.shape[0]:
            n2 = (xt * xt).sum(dim=1)
            if normalize:
                bad = bool((n2 <= 0).any())
            else:
                bad = bool(((n2 - 1.0).abs() > self._NORM_TOL).any())
            if bad:
                raise ValueError("the exact-L2 'HNSW' index ranks by inner product and therefore needs unit-norm "
                                 "rows: add through FAISSIndex.add (which normalises, as the reference does) and "
                                 "do not add zero vectors")
        super().add(xt if self._dp == self.d else xt[:, :self.d], normalize=normalize)

    def _query_sqnorms(self, x, normalize: bool):
        torch = self._torch
        if isinstance(x, torch.Tensor):
            n2 = (x.detach().to(torch.float32) ** 2).sum(dim=1)
            n2 = (n2 > 0).to(torch.float32) if normalize else n2
            return n2
        a = np.asarray(x, dtype=np.float32)
        n2 = np.einsum("ij,ij->i", a, a, dtype=np.float32)
        return (n2 > 0).astype(np.float32) if normalize else n2

    def search(self, x, k: int, *, normalize: bool = False, nprobe: int = 0, return_device: bool = False):
        S, I = super().search(x, k, normalize=normalize, return_device=return_device)
        q2 = self._query_sqnorms(x, normalize)
        if return_device:
            torch = self._torch
            q2 = q2.to(S.device) if isinstance(q2, torch.Tensor) else torch.from_numpy(q2).to(S.device)
            missing = S <= -3.0e38      # unfilled slot (-FLT_MAX); its label may already be id_map[-1]
            D = torch.clamp(q2[:, None] + 1.0 - 2.0 * torch.where(missing, torch.zeros_like(S), S), min=0.0)
            return torch.where(missing, torch.full_like(D, 3.4028234663852886e38), D), I
        if not isinstance(q2, np.ndarray):
            q2 = q2.cpu().numpy()
        missing = S <= np.float32(-3.0e38)   # unfilled slot (-FLT_MAX); its label may already be id_map[-1]
        D = np.maximum(q2[:, None] + np.float32(1.0) - np.float32(2.0) * np.where(missing, np.float32(0.0), S),
                       np.float32(0.0)).astype(np.float32)
        D[missing] = np.float32(3.4028234663852886e38)
        return D, I


_NATIVE_MAGIC = b"B2RIDX01"


class FAISSIndex:
    """Reference-compatible wrapper (faiss_retrieval.py:14-256)."""

    verbose = True  # the reference prints progress lines; set False to silence them

    def __init__(self, dimension: int, index_type: str = 'IVF', nlist: int = 100, nprobe: int = 10,
                 use_gpu: bool = False, *, device=None, pq_m: int = 8):
        self.dimension = dimension
        self.index_type = index_type
        self.nlist = nlist
        self.nprobe = nprobe
        self.use_gpu = use_gpu
        self._device = device
        self._pq_m = pq_m
        self.index = None
        self.id_map: List = []
        self._ids_all_int = True
        self._create_index()

    def _say(self, msg: str) -> None:
        if self.verbose:
            print(msg)

    def _create_index(self) -> None:
        kind = self.index_type
        if kind == 'Flat':
            self.index = IndexFlatIP(self.dimension, device=self._device)
        elif kind in ('IVF', 'IVFPQ'):
            from . import ivf  # noqa: WPS433 (kept separate: optional index families)
            self.index = ivf.create(self, kind)
        elif kind == 'HNSW':
            self.index = IndexHNSWFlat(self.dimension, 32, device=self._device)   # M = 32, efC 40, efS 16 (:67-70)
        else:
            raise ValueError(f"Unknown index type: {self.index_type}")
        self._say(f"Created {self.index_type} index with dimension {self.dimension}")

    # ---------------------------------------------------------------- train
    def train(self, embeddings) -> None:
        if self.index.is_trained:
            return
        self._say(f"Training index on {len(embeddings)} samples...")
        t0 = time.time()
        self.index.train(embeddings)
        self._say(f"Index trained in {time.time() - t0:.2f}s")

    # ------------------------------------------------------------------ add
    def add(self, embeddings, ad_ids: Optional[List] = None) -> None:
        # the reference trains on the raw input before normalising it (faiss_retrieval.py:107-115)
        if not self.index.is_trained:
            self.train(embeddings)
        n = len(embeddings)
        self._say(f"Adding {n} embeddings to index...")
        t0 = time.time()
        self.index.add(embeddings, normalize=True)
        if ad_ids is None:
            start = len(self.id_map)
            ad_ids = range(start, start + n)
        else:
            ad_ids = list(ad_ids)
            if self._ids_all_int:
                self._ids_all_int = all(isinstance(a, (int, np.integer)) for a in ad_ids)
        self.id_map.extend(ad_ids)
        self._sync_ids()
        self._say(f"Added embeddings in {time.time() - t0:.2f}s")
        self._say(f"Total index size: {self.index.ntotal}")

    def _sync_ids(self) -> None:
        """Mirror id_map on the device when it is a plain int list of the right length."""
        if self._ids_all_int and len(self.id_map) == self.index.ntotal and len(self.id_map) > 0:
            try:
                arr = np.asarray(self.id_map, dtype=np.int64)
            except (OverflowError, ValueError, TypeError):
                self._ids_all_int = False
                self.index.set_ids(None)
                return
            self.index.set_ids(arr)
        else:
            self.index.set_ids(None)

    # --------------------------------------------------------------- search
    def search(self, query_embeddings, k: int = 100, return_distances: bool = True):
        t0 = time.time()
        nprobe = self.nprobe if hasattr(self.index, 'nprobe') else 0
        if nprobe:
            self.index.nprobe = self.nprobe
        distances, labels = self.index.search(query_embeddings, k, normalize=True, nprobe=nprobe)
        if self.index._ids_set:
            ad_ids = labels  # already id_map[label] (device gather)
        else:
            # non-integer ids: host gather with python's negative-index wrap (id_map[-1])
            table = np.empty(len(self.id_map), dtype=object)
            table[:] = self.id_map
            ad_ids = table[labels]
            try:
                ad_ids = np.array(ad_ids.tolist())
            except Exception:
                pass
        ms = (time.time() - t0) * 1000
        self._say(f"Search completed in {ms:.2f}ms for {len(query_embeddings)} queries")
        if return_distances:
            return ad_ids, distances
        return ad_ids

    def batch_search(self, query_embeddings, k: int = 100, batch_size: int = 1000):
        ids_parts, dist_parts = [], []
        for lo in range(0, len(query_embeddings), batch_size):
            ids, dist = self.search(query_embeddings[lo:lo + batch_size], k)
            ids_parts.append(ids)
            dist_parts.append(dist)
        return np.vstack(ids_parts), np.vstack(dist_parts)

    # ------------------------------------------------------------ save/load
    def save(self, filepath: str, *, format: str = "faiss") -> None:
        """Index file + the reference's pickled `.metadata` side-car (same keys as
        faiss_retrieval.py:209-219).  `format="faiss"` (default) writes the `faiss.write_index` layout
        (faiss_io.py) so the file is interchangeable with the reference's; `format="native"` writes
        this package's own container.  'HNSW' always uses the native container: a faiss `IHNf` file carries
        the neighbour graph, which the exact-scan stand-in never builds (reading one works, see `load`)."""
        Path(filepath).parent.mkdir(parents=True, exist_ok=True)
        if format not in ("faiss", "native"):
            raise ValueError(f"unknown index file format {format!r}")
        if self.index_type == 'HNSW':
            format = "native"
        if format == "faiss":
            from . import faiss_io
            self.index.nprobe = self.nprobe
            faiss_io.write_index(self.index, filepath)
        else:
            self._save_native(filepath)
        with open(filepath + '.metadata', 'wb') as f:
            pickle.dump({'dimension': self.dimension, 'index_type': self.index_type, 'nlist': self.nlist,
                         'nprobe': self.nprobe, 'id_map': list(self.id_map)}, f)
        self._say(f"Index saved to {filepath}")

    def _save_native(self, filepath: str) -> None:
        """`b2r_index_save` (csrc/persist.cu): the C-ABI container any host language can write and read."""
        idx = self.index
        torch = idx._torch
        with torch.cuda.device(idx.device):
            _lib.check(idx._lib.b2r_index_save(idx._h, os.fsencode(filepath), _stream_ptr(torch, idx.device)))

    def _save_native_py(self, filepath: str) -> None:   # round-1 container (kept for its reader's test)
        state = self.index.state_dict() if hasattr(self.index, "state_dict") else {}
        n = self.index.ntotal
        if n and getattr(self.index, "stores_vectors", True):
            vecs = self.index.reconstruct_n(0, n).cpu().numpy()
        else:   # empty, or IVF-PQ (codes travel in `state`)
            vecs = np.zeros((0, self.dimension), np.float32)
        with open(filepath, "wb") as f:
            f.write(_NATIVE_MAGIC)
            blob = pickle.dumps({"index_type": self.index_type, "dimension": self.dimension,
                                 "ntotal": len(vecs), "state": state, "pq_m": self._pq_m}, protocol=4)
            f.write(struct.pack("<Q", len(blob)))
            f.write(blob)
            f.write(vecs.astype(np.float32, copy=False).tobytes())

    def load(self, filepath: str) -> None:
        """Reads either layout (sniffed from the first bytes): a `faiss.write_index` file of the
        reference (Flat / IVFFlat / IVFPQ) or this package's native container."""
        from . import faiss_io
        with open(filepath + '.metadata', 'rb') as f:
            meta = pickle.load(f)
        self.dimension = meta['dimension']
        self.index_type = meta['index_type']
        self.nlist = meta['nlist']
        self.nprobe = meta['nprobe']
        layout = faiss_io.sniff(filepath)
        if layout == "faiss":
            index = faiss_io.read_index(filepath, device=self._device)
            want = {'Flat': 'IndexFlatIP', 'IVF': 'IndexIVFFlat', 'IVFPQ': 'IndexIVFPQ',
                    'HNSW': 'IndexHNSWFlat'}.get(self.index_type)
            if type(index).__name__ != want:
                raise ValueError(f"{filepath}: holds a {type(index).__name__}, metadata says {self.index_type}")
            if index.d != self.dimension:
                raise ValueError(f"{filepath}: dimension {index.d} != metadata dimension {self.dimension}")
            self.index = index
            self._pq_m = getattr(index, "pq_m", self._pq_m)
        elif layout == "native":
            self._load_native(filepath)
        elif layout == "native-py":
            self._load_native_py(filepath)
        else:
            raise ValueError(f"{filepath}: neither a faiss index file (IxFI/IwFl/IwPQ/IHNf) nor a b200 native container")
        self.id_map = list(meta['id_map'])
        self._ids_all_int = all(isinstance(a, (int, np.integer)) for a in self.id_map)
        self._sync_ids()
        self._say(f"Index loaded from {filepath}")
        self._say(f"Index size: {self.index.ntotal}")

    def _load_native(self, filepath: str) -> None:
        """`b2r_index_load`: the handle comes back fully built (quantisers, rows / codes, id map, scan format);
        it is adopted by the faiss-shaped Python object of the type the metadata names."""
        verbose, self.verbose = self.verbose, False
        try:
            self._create_index()          # right Python class / attributes for index_type; its empty handle is replaced
        finally:
            self.verbose = verbose
        idx = self.index
        torch = idx._torch
        h = C.c_void_p()
        with torch.cuda.device(idx.device):
            _lib.check(idx._lib.b2r_index_load(C.byref(h), os.fsencode(filepath), idx.device.index,
                                               _stream_ptr(torch, idx.device)))
        idx._lib.b2r_index_destroy(idx._h)
        idx._h = h
        idx._graphs = {}

    def _load_native_py(self, filepath: str) -> None:
        with open(filepath, "rb") as f:
            f.read(8)
            (blen,) = struct.unpack("<Q", f.read(8))
            head = pickle.loads(f.read(blen))
            vecs = np.frombuffer(f.read(), dtype=np.float32).reshape(head["ntotal"], head["dimension"])
        self._pq_m = head.get("pq_m", self._pq_m)
        verbose, self.verbose = self.verbose, False
        try:
            self._create_index()
        finally:
            self.verbose = verbose
        if head.get("state") and hasattr(self.index, "load_state_dict"):
            self.index.load_state_dict(head["state"])
        if len(vecs):
            self.index.add(vecs, normalize=False)  # stored rows are already normalised

    def get_stats(self) -> Dict:
        return {
            'index_type': self.index_type,
            'dimension': self.dimension,
            'num_vectors': self.index.ntotal,
            'is_trained': self.index.is_trained,
            'nlist': self.nlist if hasattr(self, 'nlist') else None,
            'nprobe': self.nprobe if hasattr(self, 'nprobe') else None,
        }


class TwoStageRetriever:
    """Stage 1 on the B200 kernels, stage 2 = whatever ranker module the caller passes
    (reference: faiss_retrieval.py:259-369)."""

    def __init__(self, two_tower_model, transformer_ranker, faiss_index: FAISSIndex, device: str = None):
        import torch
        if device is None:
            device = 'cuda' if torch.cuda.is_available() else 'cpu'
        self.two_tower_model = two_tower_model.to(device).eval()
        self.transformer_ranker = transformer_ranker.to(device).eval() if transformer_ranker is not None else None
        self.faiss_index = faiss_index
        self.device = device

    def retrieve_and_rank(self, user_categorical, user_numerical, stage1_k: int = 500, stage2_k: int = 10,
                          ad_features_lookup: Dict = None) -> Tuple[List, List]:
        import torch
        say = self.faiss_index._say
        with torch.no_grad():
            say("\n=== Stage 1: Candidate Generation ===")
            t0 = time.time()
            user_emb = self.two_tower_model.get_user_embeddings(user_categorical.to(self.device),
                                                                user_numerical.to(self.device))
            # the embedding stays on the device: no .cpu().numpy() round trip (faiss_retrieval.py:316)
            candidate_ids, distances = self.faiss_index.search(user_emb, k=stage1_k)
            stage1_ms = (time.time() - t0) * 1000
            say(f"Stage 1 completed in {stage1_ms:.2f}ms")
            say(f"Retrieved {stage1_k} candidates")

            say("\n=== Stage 2: Transformer Ranking ===")
            t1 = time.time()
            if ad_features_lookup is None or self.transformer_ranker is None:
                say("Warning: No ad features provided, skipping stage 2")
                return candidate_ids[0].tolist(), distances[0].tolist()
            batch_user_cat = user_categorical.repeat(stage1_k, 1).to(self.device)
            batch_user_num = user_numerical.repeat(stage1_k, 1).to(self.device)
            # the reference scores an all-zero placeholder here (faiss_retrieval.py:345)
            batch_ad_cat = torch.zeros(stage1_k, 20, dtype=torch.long, device=self.device)
            predictions = self.transformer_ranker(batch_user_cat, batch_ad_cat, batch_user_num)
            ctr = torch.sigmoid(predictions['ctr']).cpu().numpy()
            order = np.argsort(ctr)[::-1][:stage2_k]
            final_ids = candidate_ids[0][order].tolist()
            final_scores = ctr[order].tolist()
            stage2_ms = (time.time() - t1) * 1000
            say(f"Stage 2 completed in {stage2_ms:.2f}ms")
            say(f"Final {stage2_k} ads selected")
            say(f"\n=== Total Time: {stage1_ms + stage2_ms:.2f}ms ===")
            return final_ids, final_scores


def benchmark_faiss_index(dimension: int = 256, num_vectors: int = 1000000, num_queries: int = 100, k: int = 100):
    """Reference surface (faiss_retrieval.py:372-437): time `add` and one `search` call for each index
    family on standard-normal data and return {index_type: {add_time, search_time_ms, per_query_ms}}."""
    say = print if FAISSIndex.verbose else (lambda *a, **kw: None)
    say("\n=== Benchmarking FAISS Indices ===")
    say(f"Vectors: {num_vectors}, Queries: {num_queries}, k: {k}, dim: {dimension}\n")
    rng = np.random.default_rng()
    vectors = rng.standard_normal((num_vectors, dimension), dtype=np.float32)
    queries = rng.standard_normal((num_queries, dimension), dtype=np.float32)
    results = {}
    for index_type, config in (('Flat', {}), ('IVF', {'nlist': 100, 'nprobe': 10}),
                               ('IVFPQ', {'nlist': 100, 'nprobe': 10}), ('HNSW', {})):
        say(f"\nTesting {index_type} index...")
        index = FAISSIndex(dimension=dimension, index_type=index_type, **config)
        t0 = time.time()
        index.add(vectors)
        add_time = time.time() - t0
        t0 = time.time()
        index.search(queries, k=k)
        search_ms = (time.time() - t0) * 1000
        results[index_type] = {'add_time': add_time, 'search_time_ms': search_ms,
                               'per_query_ms': search_ms / max(num_queries, 1)}
        say(f"  Add time: {add_time:.2f}s")
        say(f"  Search time: {search_ms:.2f}ms ({search_ms / max(num_queries, 1):.2f}ms per query)")
    say("\n=== Benchmark Summary ===")
    for index_type, m in results.items():
        say(f"{index_type:10s}: {m['per_query_ms']:.2f}ms per query")
    return results
