"""Reader / writer for the faiss index file layout used by `faiss.write_index` / `faiss.read_index`
(the reference's persistence format, faiss_retrieval.py:203-206 and :221-226) — SURVEY.md §8(f) row 1.

Covers exactly the three index families the hot path builds: `IndexFlatIP` (fourcc ``IxFI``; ``IxF2`` /
``IxFl`` are accepted on read), `IndexIVFFlat` (``IwFl``) and `IndexIVFPQ` (``IwPQ``), each with an
`ArrayInvertedLists` (``ilar``, ``full`` or ``sprs`` size table) and a flat coarse quantiser.  An
`IndexHNSWFlat` file (``IHNf``, faiss_retrieval.py:65-70) can be READ: the neighbour graph is skipped and
the vectors of its flat storage go into the exact-L2 stand-in (`faiss_retrieval.IndexHNSWFlat`); it is never
written, because that index builds no graph (`FAISSIndex.save` uses the native container for 'HNSW').

The layout below is restated from faiss's published serialiser (faiss/impl/index_write.cpp, v1.7.x —
the dependency is pinned only as `faiss-cpu>=1.7.4` in the reference's requirements.txt and is not
installable in this image).  STATUS: round-trips with itself bit-for-bit; **not validated against a
file produced by a real faiss** — `tests/test_oracle_vs_faiss.py` does that when faiss is importable.

    index header   : i32 d | i64 ntotal | i64 1<<20 | i64 1<<20 | u8 is_trained | i32 metric [| f32 metric_arg if metric>1]
    IxFI           : fourcc | header | u64 n_floats | f32[n_floats]
    ivf header     : header | u64 nlist | u64 nprobe | <quantizer index> | u8 direct_map_type | u64 len | i64[len]
    IwFl           : fourcc | ivf header | invlists
    IwPQ           : fourcc | ivf header | u8 by_residual | u64 code_size | u64 d | u64 M | u64 nbits
                     | u64 n_floats | f32[M*ksub*dsub] | invlists
    IHNf (read)    : fourcc | header | vec<f64> assign_probas | vec<i32> cum_nneighbor_per_level | vec<i32> levels
                     | vec<u64> offsets | vec<i32> neighbors | i32 entry_point | i32 max_level | i32 efConstruction
                     | i32 efSearch | i32 upper_beam | <storage index: IxF2>          (vec<T> = u64 n | T[n])
    invlists ilar  : fourcc | u64 nlist | u64 code_size | fourcc full/sprs | u64 len | u64[len] sizes
                     | per non-empty list: u8[n*code_size] codes, i64[n] ids

The two layers are separate on purpose: `parse` / `serialize` are pure numpy (CPU-testable, no GPU),
`read_index` / `write_index` move a parsed description into / out of the device index objects.
"""
from __future__ import annotations

import io
import struct
from typing import BinaryIO, Dict

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
_DUMMY = 1 << 20


def _fourcc(s: str) -> bytes:
    assert len(s) == 4
    return s.encode("ascii")


class FaissFormatError(ValueError):
    pass


# --------------------------------------------------------------------------- low-level reader
class _Reader:
    def __init__(self, f: BinaryIO):
        self.f = f

    def raw(self, n: int) -> bytes:
        b = self.f.read(n)
        if len(b) != n:
            raise FaissFormatError(f"truncated file: wanted {n} bytes, got {len(b)}")
        return b

    def unpack(self, fmt: str):
        (v,) = struct.unpack("<" + fmt, self.raw(struct.calcsize("<" + fmt)))
        return v

    def fourcc(self) -> str:
        return self.raw(4).decode("ascii", errors="replace")

    def array(self, dtype, count: int) -> np.ndarray:
        dt = np.dtype(dtype)
        return np.frombuffer(self.raw(count * dt.itemsize), dtype=dt).copy()

    def vector(self, dtype) -> np.ndarray:
        n = self.unpack("Q")
        if n > (1 << 40):
            raise FaissFormatError(f"implausible vector length {n}")
        return self.array(dtype, n)


def _read_header(r: _Reader) -> Dict:
    d = r.unpack("i")
    ntotal = r.unpack("q")
    r.unpack("q")
    r.unpack("q")
    trained = bool(r.unpack("B"))
    metric = r.unpack("i")
    metric_arg = r.unpack("f") if metric > 1 else 0.0
    if d <= 0 or ntotal < 0:
        raise FaissFormatError(f"bad header d={d} ntotal={ntotal}")
    return {"d": d, "ntotal": ntotal, "is_trained": trained, "metric": metric, "metric_arg": metric_arg}


def _read_invlists(r: _Reader) -> Dict:
    tag = r.fourcc()
    if tag == "il00":
        return {"nlist": 0, "code_size": 0, "sizes": np.zeros(0, np.int64), "codes": [], "ids": []}
    if tag != "ilar":
        raise FaissFormatError(f"unsupported inverted-list container {tag!r} (only ArrayInvertedLists 'ilar')")
    nlist = r.unpack("Q")
    code_size = r.unpack("Q")
    ltype = r.fourcc()
    table = r.vector(np.uint64)
    sizes = np.zeros(nlist, np.int64)
    if ltype == "full":
        if len(table) != nlist:
            raise FaissFormatError("'full' size table length != nlist")
        sizes[:] = table.astype(np.int64)
    elif ltype == "sprs":
        if len(table) % 2:
            raise FaissFormatError("'sprs' size table must hold (list, size) pairs")
        sizes[table[0::2].astype(np.int64)] = table[1::2].astype(np.int64)
    else:
        raise FaissFormatError(f"unknown list size table {ltype!r}")
    codes, ids = [], []
    for n in sizes:
        n = int(n)
        if n:
            codes.append(r.array(np.uint8, n * code_size).reshape(n, code_size))
            ids.append(r.array(np.int64, n))
        else:
            codes.append(np.zeros((0, code_size), np.uint8))
            ids.append(np.zeros(0, np.int64))
    return {"nlist": nlist, "code_size": code_size, "sizes": sizes, "codes": codes, "ids": ids}


def _read_index(r: _Reader) -> Dict:
    tag = r.fourcc()
    if tag in ("IxFI", "IxF2", "IxFl"):
        h = _read_header(r)
        xb = r.vector(np.float32)
        if len(xb) != h["ntotal"] * h["d"]:
            raise FaissFormatError(f"flat payload holds {len(xb)} floats, header says {h['ntotal']}x{h['d']}")
        h.update(kind="Flat", xb=xb.reshape(h["ntotal"], h["d"]))
        return h
    if tag in ("IwFl", "IwPQ"):
        h = _read_header(r)
        h["nlist"] = r.unpack("Q")
        h["nprobe"] = r.unpack("Q")
        h["quantizer"] = _read_index(r)
        dm_type = r.unpack("B")
        dm = r.vector(np.int64)
        if dm_type == 2:
            raise FaissFormatError("hashtable direct maps are not supported")
        h["direct_map"] = dm
        if tag == "IwPQ":
            h["by_residual"] = bool(r.unpack("B"))
            h["code_size"] = r.unpack("Q")
            pd, pm, pbits = r.unpack("Q"), r.unpack("Q"), r.unpack("Q")
            cent = r.vector(np.float32)
            if pd != h["d"] or pm == 0 or pd % pm or len(cent) != pm * (1 << pbits) * (pd // pm):
                raise FaissFormatError("inconsistent ProductQuantizer block")
            h["pq"] = {"d": pd, "M": pm, "nbits": pbits, "centroids": cent.reshape(pm, 1 << pbits, pd // pm)}
            h["kind"] = "IVFPQ"
        else:
            h["kind"] = "IVF"
        h["invlists"] = _read_invlists(r)
        if h["invlists"]["nlist"] not in (0, h["nlist"]):
            raise FaissFormatError("invlists nlist != index nlist")
        return h
    if tag == "IHNf":
        h = _read_header(r)
        r.vector(np.float64)                       # assign_probas
        cum = r.vector(np.int32)                   # cum_nneighbor_per_level = [0, 2M, 3M, ...]
        levels = r.vector(np.int32)
        r.vector(np.uint64)                        # offsets
        r.vector(np.int32)                         # neighbors (the graph: not needed by an exact scan)
        r.unpack("i")                              # entry_point
        r.unpack("i")                              # max_level
        ef_c, ef_s = r.unpack("i"), r.unpack("i")
        r.unpack("i")                              # upper_beam
        storage = _read_index(r)
        if storage["kind"] != "Flat" or storage["d"] != h["d"] or storage["ntotal"] != h["ntotal"]:
            raise FaissFormatError("IHNf storage must be a flat index of the same shape")
        if len(levels) != h["ntotal"]:
            raise FaissFormatError("IHNf level table length != ntotal")
        h.update(kind="HNSW", xb=storage["xb"], storage_metric=storage["metric"],
                 M=int(cum[1]) // 2 if len(cum) > 1 else 32, efConstruction=ef_c, efSearch=ef_s)
        return h
    raise FaissFormatError(f"unsupported index type fourcc {tag!r} (supported: IxFI/IxF2/IxFl, IwFl, IwPQ, IHNf)")


def parse(data) -> Dict:
    """bytes / file object -> nested description (numpy arrays, python scalars)."""
    f = io.BytesIO(data) if isinstance(data, (bytes, bytearray, memoryview)) else data
    return _read_index(_Reader(f))


# --------------------------------------------------------------------------- low-level writer
def _w_header(out: BinaryIO, d: int, ntotal: int, trained: bool, metric: int, metric_arg: float = 0.0) -> None:
    out.write(struct.pack("<iqqqBi", d, ntotal, _DUMMY, _DUMMY, 1 if trained else 0, metric))
    if metric > 1:
        out.write(struct.pack("<f", metric_arg))


def _w_vector(out: BinaryIO, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a)
    out.write(struct.pack("<Q", a.size))
    out.write(a.tobytes())


def _w_flat(out: BinaryIO, desc: Dict) -> None:
    xb = np.ascontiguousarray(desc["xb"], dtype=np.float32)
    metric = desc.get("metric", METRIC_INNER_PRODUCT)
    out.write(_fourcc("IxFI" if metric == METRIC_INNER_PRODUCT else "IxF2" if metric == METRIC_L2 else "IxFl"))
    _w_header(out, desc["d"], xb.shape[0], True, metric, desc.get("metric_arg", 0.0))
    _w_vector(out, xb.reshape(-1))


def _w_invlists(out: BinaryIO, il: Dict) -> None:
    nlist, code_size = int(il["nlist"]), int(il["code_size"])
    sizes = np.asarray([len(i) for i in il["ids"]], dtype=np.uint64)
    out.write(_fourcc("ilar"))
    out.write(struct.pack("<QQ", nlist, code_size))
    non0 = np.flatnonzero(sizes)
    if len(non0) > nlist // 2:
        out.write(_fourcc("full"))
        _w_vector(out, sizes)
    else:
        out.write(_fourcc("sprs"))
        pairs = np.empty(2 * len(non0), np.uint64)
        pairs[0::2] = non0
        pairs[1::2] = sizes[non0]
        _w_vector(out, pairs)
    for c, i in zip(il["codes"], il["ids"]):
        if len(i):
            c = np.ascontiguousarray(c, dtype=np.uint8)
            assert c.size == len(i) * code_size
            out.write(c.tobytes())
            out.write(np.ascontiguousarray(i, dtype=np.int64).tobytes())


def _w_index(out: BinaryIO, desc: Dict) -> None:
    kind = desc["kind"]
    if kind == "Flat":
        return _w_flat(out, desc)
    if kind == "HNSW":
        raise FaissFormatError("an IHNf file needs the neighbour graph, which the exact-scan 'HNSW' index does not "
                               "build; use the native container")
    if kind not in ("IVF", "IVFPQ"):
        raise FaissFormatError(f"cannot serialise kind {kind!r}")
    out.write(_fourcc("IwFl" if kind == "IVF" else "IwPQ"))
    _w_header(out, desc["d"], desc["ntotal"], desc["is_trained"], desc["metric"])
    out.write(struct.pack("<QQ", desc["nlist"], desc.get("nprobe", 1)))
    _w_index(out, desc["quantizer"])
    out.write(struct.pack("<B", 0))                      # DirectMap::NoMap
    _w_vector(out, np.zeros(0, np.int64))
    if kind == "IVFPQ":
        pq = desc["pq"]
        out.write(struct.pack("<BQ", 1 if desc.get("by_residual", True) else 0, desc["code_size"]))
        out.write(struct.pack("<QQQ", pq["d"], pq["M"], pq["nbits"]))
        _w_vector(out, np.ascontiguousarray(pq["centroids"], dtype=np.float32).reshape(-1))
    _w_invlists(out, desc["invlists"])


def serialize(desc: Dict) -> bytes:
    buf = io.BytesIO()
    _w_index(buf, desc)
    return buf.getvalue()


# --------------------------------------------------------------------------- description helpers
def flat_desc(xb: np.ndarray, metric: int = METRIC_INNER_PRODUCT) -> Dict:
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    return {"kind": "Flat", "d": xb.shape[1], "ntotal": xb.shape[0], "is_trained": True, "metric": metric, "xb": xb}


def invlists_from_assignment(payload: np.ndarray, lists: np.ndarray, nlist: int) -> Dict:
    """payload uint8 [n, code_size] in label order + list id per label -> ArrayInvertedLists description
    (within a list, entries keep ascending label order — what sequential faiss `add` calls produce)."""
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    lists = np.asarray(lists, dtype=np.int64)
    order = np.argsort(lists, kind="stable")
    sizes = np.bincount(lists, minlength=nlist).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    codes = [payload[order[offs[l]:offs[l + 1]]] for l in range(nlist)]
    ids = [order[offs[l]:offs[l + 1]].astype(np.int64) for l in range(nlist)]
    return {"nlist": nlist, "code_size": payload.shape[1], "sizes": sizes, "codes": codes, "ids": ids}


def by_label(il: Dict, ntotal: int):
    """ArrayInvertedLists description -> (payload uint8 [n, code_size] in label order, list id per label).
    Labels must be exactly 0..n-1 (sequential `add`, the only way the reference fills an index)."""
    ids = np.concatenate(il["ids"]) if il["ids"] else np.zeros(0, np.int64)
    if len(ids) != ntotal:
        raise FaissFormatError(f"inverted lists hold {len(ids)} entries, header says ntotal={ntotal}")
    if ntotal and not np.array_equal(np.sort(ids), np.arange(ntotal)):
        raise FaissFormatError("labels are not 0..ntotal-1 (add_with_ids files are not supported: the reference "
                               "keeps its own id_map and never uses them)")
    codes = np.concatenate(il["codes"]) if il["codes"] else np.zeros((0, il["code_size"]), np.uint8)
    lists = np.repeat(np.arange(il["nlist"], dtype=np.int64), il["sizes"])
    payload = np.empty_like(codes)
    payload[ids] = codes
    out_lists = np.empty(ntotal, np.int64)
    out_lists[ids] = lists
    return payload, out_lists


# --------------------------------------------------------------------------- device objects <-> files
def describe(index) -> Dict:
    """One of this package's device indexes -> faiss-format description (copies the corpus to the host)."""
    from . import _lib
    d, n = index.d, index.ntotal
    if type(index).__name__ == "IndexHNSWFlat":
        raise FaissFormatError("an IHNf file needs the neighbour graph, which the exact-scan 'HNSW' index does not "
                               "build; use the native container (FAISSIndex.save does)")
    if index.kind == _lib.KIND_FLAT:
        xb = index.reconstruct_n(0, n).cpu().numpy() if n else np.zeros((0, d), np.float32)
        return flat_desc(xb, index.metric)
    desc = {"d": d, "ntotal": n, "is_trained": bool(index.is_trained), "metric": index.metric,
            "nlist": index.nlist, "nprobe": int(index.nprobe)}
    cent = index.export_centroids() if index.is_trained else np.zeros((0, d), np.float32)
    desc["quantizer"] = flat_desc(cent, METRIC_INNER_PRODUCT)   # faiss_retrieval.py:50,57 build IndexFlatIP quantisers
    lists = index.lists_by_label() if n else np.zeros(0, np.int64)
    if index.kind == _lib.KIND_IVF_FLAT:
        desc["kind"] = "IVF"
        xb = index.reconstruct_n(0, n).cpu().numpy() if n else np.zeros((0, d), np.float32)
        payload = np.ascontiguousarray(xb, dtype=np.float32).view(np.uint8).reshape(n, d * 4)
    else:
        desc["kind"] = "IVFPQ"
        m = index.pq_m
        cb = index.export_codebooks() if index.is_trained else np.zeros((m, 256, d // m), np.float32)
        desc.update(by_residual=True, code_size=m, pq={"d": d, "M": m, "nbits": 8, "centroids": cb})
        payload = index.codes_by_label() if n else np.zeros((0, m), np.uint8)
    desc["invlists"] = invlists_from_assignment(payload, lists, index.nlist)
    return desc


def write_index(index, path: str) -> None:
    with open(path, "wb") as f:
        _w_index(f, describe(index))


def _keep_fp16_scan_copy(index, rows) -> None:
    """A file holds the rows `add` stored - for the reference's wrapper always L2-normalised (faiss_retrieval.py:115).
    Re-adding them verbatim (normalize=False) must not silently demote the 16-bit scan copy from fp16 to bf16
    (wider rescore windows, more candidates than the index that was saved): unit-norm rows keep fp16."""
    x = np.asarray(rows, dtype=np.float32)
    if x.size == 0:
        return
    n2 = np.einsum("ij,ij->i", x, x)
    if np.all(np.abs(n2 - 1.0) < 1e-3):
        index.set_param("scan_dtype", 1)


def build(desc: Dict, device=None):
    """faiss-format description -> a device index of this package."""
    from .faiss_retrieval import IndexFlatIP, IndexHNSWFlat
    from . import ivf
    kind, d = desc["kind"], desc["d"]
    if kind == "HNSW":
        if desc["metric"] != METRIC_L2 or desc["storage_metric"] != METRIC_L2:
            raise FaissFormatError("IndexHNSWFlat must be METRIC_L2 (faiss_retrieval.py:68 passes no metric)")
        index = IndexHNSWFlat(d, desc["M"], device=device)
        index.hnsw.efConstruction, index.hnsw.efSearch = desc["efConstruction"], desc["efSearch"]
        if desc["ntotal"]:
            _keep_fp16_scan_copy(index, desc["xb"])
            index.add(desc["xb"], normalize=False)      # raises unless the stored rows have unit norm
        return index
    if kind == "Flat":
        if desc["metric"] != METRIC_INNER_PRODUCT:
            raise FaissFormatError("only inner-product flat indexes are on the hot path (IndexFlatIP)")
        index = IndexFlatIP(d, device=device)
        if desc["ntotal"]:
            _keep_fp16_scan_copy(index, desc["xb"])
            index.add(desc["xb"], normalize=False)
        return index
    quant = desc["quantizer"]
    if quant["kind"] != "Flat":
        raise FaissFormatError("coarse quantiser must be a flat index")
    if kind == "IVF":
        if desc["metric"] != METRIC_INNER_PRODUCT:
            raise FaissFormatError("IndexIVFFlat must be METRIC_INNER_PRODUCT (faiss_retrieval.py:52-54)")
        index = ivf.IndexIVFFlat(d, desc["nlist"], device=device)
    else:
        pq = desc["pq"]
        if pq["nbits"] != 8 or not desc.get("by_residual", True) or desc["metric"] != METRIC_L2:
            raise FaissFormatError("IndexIVFPQ must be 8-bit, by-residual, METRIC_L2 (faiss_retrieval.py:60-62)")
        index = ivf.IndexIVFPQ(d, desc["nlist"], pq["M"], device=device)
    index.nprobe = int(desc.get("nprobe", 1))
    if desc["is_trained"]:
        if kind == "IVFPQ":
            index.import_codebooks(desc["pq"]["centroids"])
        index.import_centroids(quant["xb"])
    if desc["ntotal"]:
        payload, lists = by_label(desc["invlists"], desc["ntotal"])
        if kind == "IVF":
            # rows are re-assigned by the imported centroids on add — identical to the file's lists
            # except on exact centroid ties (the same arg-max the file's writer ran)
            rows = payload.view(np.float32).reshape(desc["ntotal"], d)
            _keep_fp16_scan_copy(index, rows)
            index.add(rows, normalize=False)
        else:
            index.add_codes(payload, lists)
    return index


def read_index(path: str, device=None):
    with open(path, "rb") as f:
        return build(parse(f), device=device)


def sniff(path: str) -> str:
    """'native' for this package's own container, 'faiss' for a faiss fourcc, else 'unknown'."""
    with open(path, "rb") as f:
        head = f.read(8)
    if head == b"B2RIDX02":
        return "native"        # b2r_index_save container (csrc/persist.cu)
    if head == b"B2RIDX01":
        return "native-py"     # round-1 pickled container, still readable
    return "faiss" if head[:4] in (b"IxFI", b"IxF2", b"IxFl", b"IwFl", b"IwPQ", b"IHNf") else "unknown"
