"""Row-sharded Flat / IVF index over several B200s (the reference has no sharding —
one faiss.Index object, faiss_retrieval.py:42-78).

One process per GPU.  Rank r owns the contiguous corpus rows [r*N/P, (r+1)*N/P) and returns
global labels (local row + base).  Every rank searches ITS shard for the SAME query batch, then the
per-rank best-first top-k lists are exchanged and merged.  Because each shard's scores are exact fp32
(rescored) before the exchange, the merged result is identical to an unsharded search.

The exchange (NCCL over NVLink / NVSwitch):
  * every rank packs its [Q,k] result into ONE int32 buffer of 2k+1 words per query (fp32 score bits,
    int32 LOCAL labels - the base is added back after the exchange - and the query's status word):
    8 instead of 12 bytes per result, one collective instead of two;
  * "sliced" mode (Q >= 4P): an all-to-all by QUERY SLICE - rank r receives every shard's lists for
    queries [r*S, (r+1)*S) only and merges just those (P x less NVLink traffic and merge work than an
    all-gather + P redundant merges), then one all-gather spreads the merged slices so that every rank
    holds the full answer;
  * "gather" mode (small batches, latency-bound): one all-gather of the packed lists, every rank merges
    all Q queries; for Q <= 256 the whole step (local search + pack + all-gather + merge) is captured in a
    CUDA graph and replayed.
The per-query status words are OR-ed across shards by the merge kernel, so a query flagged "not provably
exact" on ANY shard is flagged in the merged result on EVERY rank; `search()` re-runs flagged queries
collectively with the thresholds the shards suggest.
"""
from __future__ import annotations

import os
import warnings
from typing import Tuple

from . import _lib

__all__ = ["shard_rows", "gather_topk", "ShardedFlatIndex", "ShardedIVFIndex"]

_RETRY_BITS = _lib.ST_TOO_FEW | _lib.ST_NEED_LOWER_TAU | _lib.ST_CAND_OVERFLOW
_MAX_RETRIES = 8
_GRAPH_MAX_Q = 256


def shard_rows(total_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; ranges tile [0, total_rows) exactly."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return rank * total_rows // world, (rank + 1) * total_rows // world


def slice_rows(nq: int, world: int) -> int:
    """Queries per rank in the sliced exchange: ceil(nq / world) (the last slices are padded)."""
    return (nq + world - 1) // world


def exchange_mode(nq: int, world: int) -> str:
    """'gather' for latency-bound batches (one collective, redundant merge), 'sliced' otherwise."""
    return "sliced" if nq >= 4 * world and nq > 128 else "gather"


def gather_topk(D_local, I_local, group=None):
    """All-gather the per-rank [Q,k] results -> ([P,Q,k] scores, [P,Q,k] labels), rank-major.
    (The unpacked two-collective form: kept for host-side tests and as the reference the packed
    exchange is checked against.)"""
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
        self._bases = None
        self._graphs = {}
        self.last_status = None
        self.last_retries = 0
        self.replayed_launches = 0

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
        self._graphs = {}
        self.local.add(x, normalize=normalize)

    # ------------------------------------------------------------------ exchange
    def _shard_bases(self, device):
        torch = self._torch
        if self._bases is None or self._bases.device != device:
            self._bases = torch.tensor([shard_rows(self.total_rows, self.world, r)[0] for r in range(self.world)],
                                       dtype=torch.int64, device=device)
        return self._bases

    # the two kernels of the exchange, behind one seam so that the host logic (layouts, collectives, slicing)
    # runs on CPU tensors under gloo with numpy stand-ins (tests/test_sharding_cpu.py)
    def _pack(self, q, q_rows, k, Dl, Il, st, out):
        torch = self._torch
        _lib.check(_lib.load().b2r_topk_pack(q, q_rows, k, Dl.data_ptr(), Il.data_ptr(),
                                             st.data_ptr() if st is not None else None, self.lo, out.data_ptr(),
                                             self._largest, int(torch.cuda.current_stream(Dl.device).cuda_stream)))

    def _merge(self, P, q, q_stride, k, packed, D_out, I_out, st_out):
        """`packed`: a tensor, or the raw device pointer (int) of the peer exchange's receive area"""
        torch = self._torch
        dev = D_out.device
        ptr = packed if isinstance(packed, int) else packed.data_ptr()
        _lib.check(_lib.load().b2r_topk_merge_packed(P, q, q_stride, k, ptr,
                                                     self._shard_bases(dev).data_ptr(), D_out.data_ptr(),
                                                     I_out.data_ptr(), st_out.data_ptr(), self._largest,
                                                     int(torch.cuda.current_stream(dev).cuda_stream)))

    # ---- peer-memory exchange (csrc/peer.cu): P2P stores over NVLink instead of the NCCL all-gather, for the
    #      latency-bound "gather" batches.  On by default for CUDA ranks of the default process group;
    #      B2R_PEER_EXCHANGE=0 or `index.peer_exchange = False` selects the NCCL all-gather.  If any rank fails to
    #      map a peer's buffer (no CUDA IPC in this container, no P2P path), ALL ranks fall back to NCCL together.
    peer_exchange = None          # None: follow the environment variable (default on)

    def _peer_enabled(self, dev) -> bool:
        if getattr(self, "_peer_broken", False) or dev.type != "cuda":
            return False
        want = self.peer_exchange if self.peer_exchange is not None else os.environ.get("B2R_PEER_EXCHANGE", "1") != "0"
        return bool(want) and self.world > 1 and getattr(self, "group", None) is None

    def _peer_ctx(self, need_bytes: int):
        """The (lazily created) exchange context of this index: one IPC-shared buffer per rank.  Collective."""
        import ctypes as C
        import torch.distributed as dist
        ctx = getattr(self, "_peer", None)
        if ctx is not None and ctx["cap"] >= need_bytes:
            return ctx
        torch = self._torch
        lib = _lib.load()
        dev = self.local.device
        if ctx is not None:
            torch.cuda.synchronize(dev)
            dist.barrier()                       # nobody may still be pushing into the buffer that goes away
            lib.b2r_peer_destroy(ctx["h"])
            self._graphs = {}
        cap = max(need_bytes, 4 << 20)
        h = C.c_void_p()
        ok, why = 1, ""
        with torch.cuda.device(dev):
            blob = (C.c_ubyte * 64)()
            if lib.b2r_peer_create(C.byref(h), self.rank, self.world, cap, dev.index) != 0 or \
                    lib.b2r_peer_handle(h, blob) != 0:
                ok, why = 0, (lib.b2r_last_error() or b"").decode()
            mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
            allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine)
            if ok and lib.b2r_peer_connect(h, bytes(allh.cpu().numpy().tobytes())) != 0:
                ok, why = 0, (lib.b2r_last_error() or b"").decode()
            agree = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)     # also: every rank has mapped every buffer before the first push
        if int(agree.item()) == 0:
            if h:
                lib.b2r_peer_destroy(h)
            self._peer, self._peer_broken = None, True
            warnings.warn("sharded search: peer-memory exchange unavailable on some rank"
                          + (f" ({why})" if why else "") + "; using the NCCL all-gather")
            return None
        self._peer = {"h": h, "cap": cap}
        return self._peer

    def _peer_allgather(self, send, nbytes: int) -> int:
        import ctypes as C
        torch = self._torch
        ctx = self._peer
        recv = C.c_void_p()
        with torch.cuda.device(send.device):
            _lib.check(_lib.load().b2r_peer_allgather(ctx["h"], send.data_ptr(), nbytes, C.byref(recv),
                                                      int(torch.cuda.current_stream(send.device).cuda_stream)))
        return int(recv.value)

    def _peer_ack(self, dev) -> None:
        torch = self._torch
        with torch.cuda.device(dev):
            _lib.check(_lib.load().b2r_peer_ack(self._peer["h"], int(torch.cuda.current_stream(dev).cuda_stream)))

    def close(self) -> None:
        """Release the captured graphs and the peer-exchange buffers (collective when the latter exist)."""
        self.release_graphs()
        ctx = getattr(self, "_peer", None)
        if ctx is not None:
            import torch.distributed as dist
            self._torch.cuda.synchronize(self.local.device)
            dist.barrier()
            _lib.load().b2r_peer_destroy(ctx["h"])
            self._peer = None

    def _exchange(self, Dl, Il, st, k: int, bufs=None):
        """Packed exchange + merge of this rank's local result.  Returns (D, I, status) [Q,k] / [Q] tensors,
        identical on every rank.  `bufs` (dict) caches the scratch tensors of one shape (graph replay)."""
        import torch.distributed as dist
        torch = self._torch
        dev = Dl.device
        P, Q = self.world, Dl.shape[0]
        W = 2 * k + 1
        bufs = {} if bufs is None else bufs

        def buf(name, shape, dtype):
            t = bufs.get(name)
            if t is None or tuple(t.shape) != tuple(shape):
                t = bufs[name] = torch.empty(shape, dtype=dtype, device=dev)
            return t

        if exchange_mode(Q, P) == "gather":
            D_out = buf("D", (Q, k), torch.float32)
            I_out = buf("I", (Q, k), torch.int64)
            st_out = buf("st", (Q,), torch.int32)
            nbytes = (Q * W * 4 + 15) // 16 * 16
            if self._peer_enabled(dev):
                # packed list -> slot [rank] of every rank's receive area by P2P stores; rank r's block starts at
                # recv + r * nbytes, so the merge's block stride is nbytes / 4 words = (nbytes / 4 / W) rows only when
                # Q * W * 4 is a multiple of 16: pad the row count of the packed buffer instead
                rows = Q
                while (rows * W * 4) % 16:
                    rows += 1
                nbytes = rows * W * 4
                ctx = getattr(self, "_peer", None)
                if not torch.cuda.is_current_stream_capturing():
                    ctx = self._peer_ctx(P * nbytes)
                if ctx is not None:
                    send = buf("send", (rows, W), torch.int32)
                    self._pack(Q, rows, k, Dl, Il, st, send)
                    recv_ptr = self._peer_allgather(send, nbytes)
                    self._merge(P, Q, rows, k, recv_ptr, D_out, I_out, st_out)
                    self._peer_ack(dev)
                    return D_out, I_out, st_out
            send = buf("send", (Q, W), torch.int32)
            recv = buf("recv", (P, Q, W), torch.int32)
            self._pack(Q, Q, k, Dl, Il, st, send)
            if dev.type == "cuda":
                dist.all_gather_into_tensor(recv, send, group=self.group)
            else:
                dist.all_gather(list(recv.unbind(0)), send, group=self.group)
            self._merge(P, Q, Q, k, recv, D_out, I_out, st_out)
            return D_out, I_out, st_out
        # ---- sliced: all-to-all by query slice -> merge my slice -> all-gather the merged slices
        S = slice_rows(Q, P)
        send = buf("send", (P * S, W), torch.int32)
        recv = buf("recv", (P, S, W), torch.int32)
        self._pack(Q, P * S, k, Dl, Il, st, send)
        dist.all_to_all_single(recv.view(P * S, W), send, group=self.group)
        # merged slice = one flat int32 buffer [scores S*k | labels (int64) S*k | status S]: the merge kernel
        # writes straight into it and ONE all-gather spreads it
        off_i = (S * k + 1) // 2 * 2
        off_s = off_i + 2 * S * k
        L = (off_s + S + 1) // 2 * 2
        flat = buf("flat", (L,), torch.int32)
        Dm = flat[:S * k].view(torch.float32).view(S, k)
        Im = flat[off_i:off_s].view(torch.int64).view(S, k)
        sm = flat[off_s:off_s + S]
        self._merge(P, S, S, k, recv, Dm, Im, sm)
        allf = buf("allf", (P, L), torch.int32)
        if dev.type == "cuda":
            dist.all_gather_into_tensor(allf, flat, group=self.group)
        else:
            dist.all_gather(list(allf.unbind(0)), flat, group=self.group)
        D_full = buf("D", (P * S, k), torch.float32)
        I_full = buf("I", (P * S, k), torch.int64)
        st_full = buf("st", (P * S,), torch.int32)
        D_full.view(P, S * k).copy_(allf[:, :S * k].view(torch.float32))
        I_full.view(P, S * k).copy_(allf[:, off_i:off_s].view(torch.int64))
        st_full.view(P, S).copy_(allf[:, off_s:off_s + S])
        return D_full[:Q], I_full[:Q], st_full[:Q]

    def _search_eager(self, qt, k, normalize, tau=None, **local_kw):
        Dl, Il, st, tr = self.local.search_device(qt, k, normalize=normalize, tau=tau, **local_kw)
        if self.world == 1:
            return Dl, Il, st, tr
        D, I, st_all = self._exchange(Dl, Il, st, k)
        return D, I, st_all, tr

    def search_device(self, q, k: int, normalize: bool = True, tau=None, **local_kw):
        """(D [Q,k], I [Q,k], status [Q]) CUDA tensors, identical on every rank.  `status` is the OR of every
        shard's status word (0 = provably exact on all shards).  Asynchronous; no retry (see `search`).
        Batches up to 256 queries replay a CUDA graph of the whole step: their result tensors are STATIC
        (valid until the next search of that shape)."""
        torch = self._torch
        nq = len(q)
        if (self.world > 1 and tau is None and 0 < nq <= _GRAPH_MAX_Q and not os.environ.get("B2R_NO_GRAPHS")
                and isinstance(q, torch.Tensor) and q.is_cuda):
            ent = self._graph_entry(nq, int(k), bool(normalize), tuple(sorted(local_kw.items())))
            if ent:
                ent["q"].copy_(q, non_blocking=True)
                ent["graph"].replay()
                self.replayed_launches += ent["launches"]
                return ent["D"], ent["I"], ent["st"]
        D, I, st, _ = self._search_eager(q, k, normalize, tau, **local_kw)
        return D, I, st

    def release_graphs(self) -> None:
        """Drop the captured search graphs.  They hold NCCL kernels of the process group's communicator, so they
        MUST be released before `dist.destroy_process_group()` (the communicator's teardown otherwise waits for
        them forever: a 2-GPU bench at batch 64 hung at exit that way)."""
        if self._graphs:
            self._graphs = {}
            import gc
            gc.collect()
            self._torch.cuda.synchronize(self.local.device)

    def __del__(self):
        try:
            self._graphs = {}
        except Exception:
            pass

    # ---- small batches: capture local search + pack + all-gather + merge once, replay afterwards
    def _graph_entry(self, nq, k, normalize, kw):
        key = (nq, k, normalize, kw)
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        torch = self._torch
        dev = self.local.device
        lib = _lib.load()
        ent = {"q": torch.zeros((nq, self.d), dtype=torch.float32, device=dev), "bufs": {}}

        def run():
            Dl, Il, st, _ = self.local.search_device(ent["q"], k, normalize=normalize, **dict(kw))
            ent["keep"] = (Dl, Il, st)
            return self._exchange(Dl, Il, st, k, ent["bufs"])

        keep_ws = self.local._ws
        try:
            with torch.cuda.device(dev):
                n0 = int(lib.b2r_debug_launch_count())
                run()                                  # eager twice: kernel attributes, NCCL channels, scratch buffers
                run()
                ent["launches"] = (int(lib.b2r_debug_launch_count()) - n0) // 2
                torch.cuda.synchronize(dev)
                # the captured search must own its workspace: the index's shared one may be re-grown (freed) later
                self.local._ws = None
                graph = torch.cuda.CUDAGraph()
                # thread_local: the NCCL watchdog thread's event queries must not invalidate the capture
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    ent["D"], ent["I"], ent["st"] = run()
                ent["ws"] = self.local._ws
                ent["graph"] = graph
        except Exception as exc:   # not capturable with this NCCL / driver: stay eager for this shape
            warnings.warn(f"sharded search: CUDA-graph capture failed ({type(exc).__name__}: {exc}); staying eager")
            ent = False
            torch.cuda.synchronize(dev)
        finally:
            self.local._ws = keep_ws
        self._graphs[key] = ent
        return ent

    def search(self, q, k: int, normalize: bool = True, **local_kw):
        """Host-result search with the COLLECTIVE retry: (D, I) numpy arrays, identical on every rank.
        Queries flagged on any shard (the status words are OR-ed across shards) are re-run on EVERY shard with
        the threshold that shard suggested, exchanged and merged again, until none is flagged."""
        torch = self._torch
        qt = self.local._to_device_f32(q, "search")
        D, I, st, tau_retry = self._search_eager(qt, k, normalize, **local_kw)
        D, I = D.clone(), I.clone()
        st_h = st.cpu().numpy().copy()           # synchronises
        retries = 0
        while (self.world > 1 and self.local._supports_retry and getattr(self.local, "_collective_retry", True)
               and retries < _MAX_RETRIES):
            bad = (st_h & _RETRY_BITS).nonzero()[0]
            if bad.size == 0:
                break
            bad_t = torch.as_tensor(bad, device=qt.device)
            tau = tau_retry.index_select(0, bad_t).contiguous()
            D2, I2, st2, tr = self._search_eager(qt.index_select(0, bad_t), k, normalize, tau, **local_kw)
            D.index_copy_(0, bad_t, D2)
            I.index_copy_(0, bad_t, I2)
            tau_retry = tau_retry.clone()
            tau_retry.index_copy_(0, bad_t, tr)
            st_h[bad] = st2.cpu().numpy()
            retries += 1
        self.last_status, self.last_retries = st_h, retries
        if (st_h != 0).any():
            warnings.warn(f"sharded search: {int((st_h != 0).sum())} queries not provably exact on some shard "
                          f"(status bits {sorted(set(int(s) for s in st_h if s))})")
        return D.cpu().numpy(), I.cpu().numpy()


class ShardedIVFIndex(ShardedFlatIndex):
    """IVF-Flat (inner product) or IVF-PQ (L2) index sharded by row ownership: the coarse
    centroids (and PQ codebooks) are REPLICATED, every rank files its own rows into its own copy of the
    inverted lists, every rank probes the same `nprobe` lists for the same queries, and the per-rank
    top-k lists take the same exchange + merge as the flat index.  With shared quantisers the merged
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
            local = ivf.IndexIVFFlat(self.d, self.nlist, device=device)
            # a flagged query of the fused list scan would need a COLLECTIVE re-run on every shard; the sharded
            # index therefore keeps the dump path, whose exact fallback is inside the kernel
            local.set_param("ivf_fused", 0)
            return local
        return ivf.IndexIVFPQ(self.d, self.nlist, self.pq_m, device=device)

    def train(self, x, src: int = 0) -> None:
        """Collective.  Rank `src` trains on ITS `x` (other ranks' `x` is ignored and may be None); the
        centroids / codebooks are then broadcast so that every rank holds bit-identical quantisers."""
        import torch.distributed as dist
        torch = self._torch
        dev = self.local.device
        self._graphs = {}
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
        self._graphs = {}
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

    def search_device(self, q, k: int, normalize: bool = True, nprobe: int = 0, tau=None):
        return super().search_device(q, k, normalize=normalize, nprobe=nprobe or self.nprobe)

    def _search_eager(self, qt, k, normalize, tau=None, **local_kw):
        local_kw.setdefault("nprobe", self.nprobe)
        return super()._search_eager(qt, k, normalize, None, **local_kw)

    def search(self, q, k: int, normalize: bool = True, nprobe: int = 0):
        return super().search(q, k, normalize=normalize, nprobe=nprobe or self.nprobe)
