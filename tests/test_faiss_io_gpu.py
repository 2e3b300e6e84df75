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
