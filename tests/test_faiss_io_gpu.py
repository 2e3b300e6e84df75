"""Index persistence in the faiss file layout (faiss_retrieval.py:203-226; faiss_io.py): every index
family round-trips through `FAISSIndex.save/load` in both layouts with bit-identical answers, the file
has the structure the layout demands, and a file authored on the host (not by our device writer) loads
and answers like the oracle."""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _clustered(n, d, ncl, seed):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((ncl, d)).astype(np.float32)
    lab = rng.integers(0, ncl, n)
    return (centres[lab] + 0.35 * rng.standard_normal((n, d))).astype(np.float32)


@pytest.fixture(scope="module")
def fr(built_lib):
    from movie_recommender_demo_b200 import faiss_retrieval
    faiss_retrieval.FAISSIndex.verbose = False
    return faiss_retrieval


@pytest.mark.parametrize("kind,fourcc", [("Flat", b"IxFI"), ("IVF", b"IwFl"), ("IVFPQ", b"IwPQ")])
def test_round_trip_both_layouts(fr, tmp_path, kind, fourcc):
    from movie_recommender_demo_b200 import faiss_io
    d, N = 64, 12000
    x = _clustered(N, d, 40, seed=3)
    q = _clustered(9, d, 40, seed=4)
    ad_ids = [1000 + 3 * i for i in range(N)]
    g = fr.FAISSIndex(d, kind, nlist=20, nprobe=5)
    g.add(x, ad_ids)
    ids, dist = g.search(q, k=60)
    results = {}
    for layout in ("faiss", "native"):
        path = str(tmp_path / f"{kind}.{layout}")
        g.save(path, format=layout)
        assert faiss_io.sniff(path) == layout
        assert set(pickle.load(open(path + ".metadata", "rb"))) == {'dimension', 'index_type', 'nlist', 'nprobe', 'id_map'}
        h = fr.FAISSIndex(d, 'Flat')
        h.load(path)
        assert h.index_type == kind and h.index.ntotal == N and h.id_map == ad_ids
        # a loaded index must search like the saved one: same 16-bit scan format (fp16 for normalised rows),
        # hence the same rescore window and candidate target
        if kind != "IVFPQ":      # IVF-PQ stores codes only: no 16-bit scan copy
            assert h.index.get_param("scan_dtype") == g.index.get_param("scan_dtype")
        results[layout] = h.search(q, k=60)
        assert np.array_equal(results[layout][0], ids)
        if kind == "IVFPQ":      # codes + codebooks travel verbatim -> same ADC scores
            assert np.allclose(results[layout][1], dist, rtol=1e-6, atol=1e-6)
        else:
            assert np.array_equal(results[layout][1], dist)
    # structure of the faiss-layout file
    path = str(tmp_path / f"{kind}.faiss")
    assert open(path, "rb").read(4) == fourcc
    desc = faiss_io.parse(open(path, "rb"))
    assert desc["d"] == d and desc["ntotal"] == N and desc["is_trained"]
    if kind == "Flat":
        assert desc["metric"] == 0
        assert np.allclose(np.linalg.norm(desc["xb"], axis=1), 1.0, atol=1e-5)       # normalised on add (:115)
    else:
        assert desc["nlist"] == 20 and desc["nprobe"] == 5 and desc["quantizer"]["ntotal"] == 20
        assert desc["metric"] == (0 if kind == "IVF" else 1)
        assert np.array_equal(desc["invlists"]["sizes"], g.index.list_sizes())
        assert desc["invlists"]["code_size"] == (4 * d if kind == "IVF" else 8)
        for l, lab in enumerate(desc["invlists"]["ids"]):
            assert (np.diff(lab) > 0).all(), f"list {l}: labels must ascend (sequential add order)"
        if kind == "IVF":       # every stored vector sits in the list of its best centroid
            cent = desc["quantizer"]["xb"]
            for l in (0, 7, 19):
                v = desc["invlists"]["codes"][l].view(np.float32).reshape(-1, d)
                if len(v):
                    s = v @ cent.T
                    assert (s.max(1) - s[:, l] <= 1e-5).all()


def test_host_authored_ivf_file_answers_like_the_oracle(fr, tmp_path):
    """A file assembled on the host from numpy (standing in for one the reference wrote with
    faiss.write_index) loads through FAISSIndex.load and agrees with the oracle on the same centroids."""
    from movie_recommender_demo_b200 import faiss_io
    from oracle.compare import compare_topk
    from oracle.flat import normalize_L2
    from oracle.ivf import OracleIndexIVFFlat
    d, N, nlist = 64, 8000, 16
    x = normalize_L2(_clustered(N, d, 32, seed=11))
    q = _clustered(12, d, 32, seed=12)
    o = OracleIndexIVFFlat(d, nlist)
    o.train(x)
    o.add(x)
    cent = np.asarray(o.centroids, dtype=np.float32)
    lists = np.argmax(x @ cent.T, axis=1)
    desc = {"kind": "IVF", "d": d, "ntotal": N, "is_trained": True, "metric": 0, "nlist": nlist, "nprobe": 1,
            "quantizer": faiss_io.flat_desc(cent),
            "invlists": faiss_io.invlists_from_assignment(x.view(np.uint8).reshape(N, 4 * d), lists, nlist)}
    path = str(tmp_path / "ref.index")
    open(path, "wb").write(faiss_io.serialize(desc))
    pickle.dump({'dimension': d, 'index_type': 'IVF', 'nlist': nlist, 'nprobe': 4, 'id_map': list(range(N))},
                open(path + ".metadata", "wb"))
    h = fr.FAISSIndex(d, 'Flat')
    h.load(path)
    assert h.index.ntotal == N and h.nprobe == 4
    assert np.array_equal(h.index.list_sizes(), np.bincount(lists, minlength=nlist))
    ids, dist = h.search(q, k=100)
    o.nprobe = 4
    rd, rid = o.search(normalize_L2(q.copy()), 100, extra=16)
    compare_topk(ids, dist, rid, rd, 100, gap_tol=1e-6)


def test_load_rejects_mismatched_metadata(fr, tmp_path):
    d = 64
    g = fr.FAISSIndex(d, 'Flat')
    g.add(_clustered(500, d, 4, seed=1))
    path = str(tmp_path / "f.index")
    g.save(path)
    meta = pickle.load(open(path + ".metadata", "rb"))
    meta['index_type'] = 'IVF'
    pickle.dump(meta, open(path + ".metadata", "wb"))
    with pytest.raises(ValueError, match="metadata says IVF"):
        fr.FAISSIndex(d, 'Flat').load(path)
    open(path, "wb").write(b"not an index")
    with pytest.raises(ValueError, match="neither a faiss index file"):
        fr.FAISSIndex(d, 'Flat').load(path)


def test_native_container_through_the_c_abi_only(fr, built_lib, tmp_path):
    """b2r_index_save / b2r_index_load called the way a non-Python host would (no wrapper, no metadata side-car):
    the loaded handle returns the same labels / mapped ids / scores as the saved one, for all three kinds."""
    import ctypes as C
    import torch
    from movie_recommender_demo_b200 import _lib
    lib = built_lib
    d, N, k = 128, 20000, 50
    x = torch.from_numpy(_clustered(N, d, 40, seed=5)).cuda()
    q = torch.from_numpy(_clustered(7, d, 40, seed=6)).cuda()
    sp = int(torch.cuda.current_stream().cuda_stream)
    for kind, nlist, pq_m, metric in ((0, 0, 0, 0), (1, 16, 0, 0), (2, 16, 16, 1)):
        h = C.c_void_p()
        _lib.check(lib.b2r_index_create(C.byref(h), kind, d, nlist, pq_m, 8 if pq_m else 0, metric, 0))
        if kind:
            _lib.check(lib.b2r_index_train(h, N, x.data_ptr(), 1234, sp))
        _lib.check(lib.b2r_index_add(h, N, x.data_ptr(), 1, sp))
        ids = (torch.arange(N, device="cuda", dtype=torch.int64) * 5 + 11)
        _lib.check(lib.b2r_index_set_ids(h, N, ids.data_ptr(), sp))

        def search(handle):
            D = torch.empty((7, k), dtype=torch.float32, device="cuda")
            I = torch.empty((7, k), dtype=torch.int64, device="cuda")
            st = torch.zeros(7, dtype=torch.int32, device="cuda")
            need = int(lib.b2r_index_search_workspace(handle, 7, k, 4))
            ws = torch.empty(max(need, 256), dtype=torch.uint8, device="cuda")
            _lib.check(lib.b2r_index_search(handle, 7, q.data_ptr(), 1, k, 4, D.data_ptr(), I.data_ptr(), st.data_ptr(),
                                            None, None, ws.data_ptr(), ws.numel(), sp))
            torch.cuda.synchronize()
            return D.cpu().numpy(), I.cpu().numpy(), st.cpu().numpy()

        D0, I0, st0 = search(h)
        path = str(tmp_path / f"kind{kind}.b2r").encode()
        _lib.check(lib.b2r_index_save(h, path, sp))
        assert open(path, "rb").read(8) == b"B2RIDX02"
        h2 = C.c_void_p()
        _lib.check(lib.b2r_index_load(C.byref(h2), path, 0, sp))
        assert lib.b2r_index_ntotal(h2) == N and lib.b2r_index_is_trained(h2) == 1
        if kind != 2:      # IVF-PQ stores codes only: no 16-bit scan copy whose format could change
            assert lib.b2r_index_get_param(h2, b"scan_dtype") == lib.b2r_index_get_param(h, b"scan_dtype")
        D1, I1, st1 = search(h2)
        assert np.array_equal(I0, I1) and (I0 % 5 == 1).all()
        assert np.array_equal(D0, D1) if kind != 2 else np.allclose(D0, D1, rtol=1e-6, atol=1e-6)
        assert (st0 == 0).all() and (st1 == 0).all()
        lib.b2r_index_destroy(h)
        lib.b2r_index_destroy(h2)
    bad = tmp_path / "bad.b2r"
    bad.write_bytes(b"not an index")
    h3 = C.c_void_p()
    assert lib.b2r_index_load(C.byref(h3), str(bad).encode(), 0, sp) != 0 and not h3.value
