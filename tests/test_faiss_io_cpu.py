"""faiss index file layout (movie_recommender_demo_b200/faiss_io.py), host-only layer: parse/serialize.
The expected bytes below are assembled by hand from the published layout (faiss/impl/index_write.cpp),
not by the code under test."""
import io
import struct

import numpy as np
import pytest

from movie_recommender_demo_b200 import faiss_io as fio


def _hdr(d, n, trained, metric):
    return struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<qq", 1 << 20, 1 << 20) + \
        struct.pack("<B", trained) + struct.pack("<i", metric)


def test_flat_ip_bytes_are_the_published_layout():
    xb = np.arange(8, dtype=np.float32).reshape(2, 4)
    want = b"IxFI" + _hdr(4, 2, 1, 0) + struct.pack("<Q", 8) + xb.tobytes()
    got = fio.serialize(fio.flat_desc(xb))
    assert got == want
    back = fio.parse(got)
    assert back["kind"] == "Flat" and back["d"] == 4 and back["ntotal"] == 2 and back["metric"] == 0
    assert np.array_equal(back["xb"], xb)


def test_ivfflat_bytes_full_and_sparse_size_tables():
    d, nlist = 4, 4
    cent = np.eye(4, dtype=np.float32)
    xb = np.array([[1, 0, 0, 0], [0, 0, 2, 0], [3, 0, 0, 0]], np.float32)
    lists = np.array([0, 2, 0])
    desc = {"kind": "IVF", "d": d, "ntotal": 3, "is_trained": True, "metric": 0, "nlist": nlist, "nprobe": 2,
            "quantizer": fio.flat_desc(cent),
            "invlists": fio.invlists_from_assignment(xb.view(np.uint8).reshape(3, 16), lists, nlist)}
    got = fio.serialize(desc)
    quant = b"IxFI" + _hdr(4, 4, 1, 0) + struct.pack("<Q", 16) + cent.tobytes()
    want = (b"IwFl" + _hdr(4, 3, 1, 0) + struct.pack("<QQ", nlist, 2) + quant
            + struct.pack("<B", 0) + struct.pack("<Q", 0)                       # direct map: NoMap, empty array
            + b"ilar" + struct.pack("<QQ", nlist, 16)
            + b"sprs" + struct.pack("<Q", 4) + struct.pack("<QQQQ", 0, 2, 2, 1)  # 2 of 4 lists non-empty -> sparse
            + xb[[0, 2]].tobytes() + struct.pack("<qq", 0, 2)
            + xb[[1]].tobytes() + struct.pack("<q", 1))
    assert got == want
    back = fio.parse(got)
    payload, l2 = fio.by_label(back["invlists"], 3)
    assert np.array_equal(payload.view(np.float32).reshape(3, 4), xb) and np.array_equal(l2, lists)
    assert back["nprobe"] == 2 and np.array_equal(back["quantizer"]["xb"], cent)

    # 3 of 4 lists non-empty -> 'full' table
    lists = np.array([0, 2, 3])
    desc["invlists"] = fio.invlists_from_assignment(xb.view(np.uint8).reshape(3, 16), lists, nlist)
    got = fio.serialize(desc)
    assert b"full" + struct.pack("<Q", 4) + struct.pack("<QQQQ", 1, 0, 1, 1) in got
    payload, l2 = fio.by_label(fio.parse(got)["invlists"], 3)
    assert np.array_equal(l2, lists)


def test_ivfpq_round_trip_and_pq_block():
    rng = np.random.default_rng(0)
    d, nlist, m, n = 16, 8, 4, 100
    cb = rng.standard_normal((m, 256, d // m)).astype(np.float32)
    codes = rng.integers(0, 256, (n, m), dtype=np.uint8)
    lists = rng.integers(0, nlist, n)
    desc = {"kind": "IVFPQ", "d": d, "ntotal": n, "is_trained": True, "metric": 1, "nlist": nlist, "nprobe": 3,
            "quantizer": fio.flat_desc(rng.standard_normal((nlist, d)).astype(np.float32)),
            "by_residual": True, "code_size": m, "pq": {"d": d, "M": m, "nbits": 8, "centroids": cb},
            "invlists": fio.invlists_from_assignment(codes, lists, nlist)}
    blob = fio.serialize(desc)
    assert blob[:4] == b"IwPQ"
    pq_block = struct.pack("<BQ", 1, m) + struct.pack("<QQQ", d, m, 8) + struct.pack("<Q", cb.size) + cb.tobytes()
    assert pq_block in blob
    back = fio.parse(blob)
    assert back["kind"] == "IVFPQ" and back["metric"] == 1 and back["by_residual"] and back["code_size"] == m
    assert np.array_equal(back["pq"]["centroids"], cb)
    c2, l2 = fio.by_label(back["invlists"], n)
    assert np.array_equal(c2, codes) and np.array_equal(l2, lists)
    assert fio.serialize(back) == blob            # idempotent


def _hnsw_file(xb, M=32, ef_c=40, ef_s=16, storage_metric=1):
    """An `IndexHNSWFlat` file authored from the published layout (header, HNSW block, flat L2 storage);
    the graph arrays hold plausible sizes and arbitrary contents - an exact scan never looks at them."""
    n, d = xb.shape
    out = io.BytesIO()
    out.write(b"IHNf")
    out.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, 1))
    def vec(a):
        a = np.ascontiguousarray(a)
        out.write(struct.pack("<Q", a.size))
        out.write(a.tobytes())
    vec(np.array([0.9, 0.09, 0.01], np.float64))                   # assign_probas
    vec(np.array([0, 2 * M, 3 * M, 4 * M], np.int32))              # cum_nneighbor_per_level
    vec(np.ones(n, np.int32))                                      # levels
    vec(np.arange(n + 1, dtype=np.uint64) * (2 * M))               # offsets
    vec(np.full(n * 2 * M, -1, np.int32))                          # neighbors
    out.write(struct.pack("<iiiii", 0, 0, ef_c, ef_s, 1))          # entry_point, max_level, efC, efS, upper_beam
    out.write(fio.serialize(fio.flat_desc(xb, storage_metric)))
    return out.getvalue()


def test_hnsw_file_is_read_as_vectors_plus_parameters():
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((37, 12)).astype(np.float32)
    desc = fio.parse(_hnsw_file(xb, M=16, ef_c=55, ef_s=21))
    assert desc["kind"] == "HNSW" and desc["metric"] == 1 and desc["storage_metric"] == 1
    assert desc["M"] == 16 and desc["efConstruction"] == 55 and desc["efSearch"] == 21
    assert np.array_equal(desc["xb"], xb)
    with pytest.raises(fio.FaissFormatError, match="neighbour graph"):
        fio.serialize(desc)                                        # never written: there is no graph to write
    with pytest.raises(fio.FaissFormatError, match="truncated"):
        fio.parse(_hnsw_file(xb)[:-3])


def test_reader_rejects_what_it_cannot_represent(tmp_path):
    with pytest.raises(fio.FaissFormatError, match="unsupported index type"):
        fio.parse(b"IHNs" + b"\0" * 64)
    good = fio.serialize(fio.flat_desc(np.ones((3, 4), np.float32)))
    with pytest.raises(fio.FaissFormatError, match="truncated"):
        fio.parse(good[:-5])
    # add_with_ids style labels
    il = fio.invlists_from_assignment(np.zeros((2, 4), np.uint8), np.array([0, 1]), 2)
    il["ids"][1] = np.array([77])
    with pytest.raises(fio.FaissFormatError, match="not 0..ntotal-1"):
        fio.by_label(il, 2)
    p = tmp_path / "x.index"
    p.write_bytes(good)
    assert fio.sniff(str(p)) == "faiss"
    p.write_bytes(b"B2RIDX01....")
    assert fio.sniff(str(p)) == "native-py"        # round-1 pickled container
    p.write_bytes(b"B2RIDX02....")
    assert fio.sniff(str(p)) == "native"           # b2r_index_save container (csrc/persist.cu)
    p.write_bytes(b"garbage!")
    assert fio.sniff(str(p)) == "unknown"
