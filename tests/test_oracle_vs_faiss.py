"""If a real `faiss` is importable (it is NOT in this image), pin the oracle against it.  Skipped otherwise —
which is why DESIGN.md §5 calls retrieval parity 'unpinned'."""
import numpy as np
import pytest

faiss = pytest.importorskip("faiss", reason="faiss-cpu is not installed in this image (no wheel, no network)")


def _data(n=5000, d=64, q=16, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, d)).astype(np.float32), rng.standard_normal((q, d)).astype(np.float32)


def test_flat_ip_matches_faiss():
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex, normalize_L2
    x, q = _data()
    o = OracleFAISSIndex(64, 'Flat')
    o.add(x)
    xs, qs = x.copy(), q.copy()
    faiss.normalize_L2(xs)
    faiss.normalize_L2(qs)
    assert np.allclose(xs, normalize_L2(x.copy()), atol=1e-7)
    idx = faiss.IndexFlatIP(64)
    idx.add(xs)
    D, I = idx.search(qs, 100)
    rid, rd = o.search(q, k=100, extra=16)
    compare_topk(I, D, rid, rd, 100, gap_tol=1e-6)


def test_ivf_flat_matches_faiss_on_shared_centroids():
    from oracle.compare import compare_topk
    from oracle.ivf import OracleIndexIVFFlat
    x, q = _data()
    faiss.normalize_L2(x)
    faiss.normalize_L2(q)
    quant = faiss.IndexFlatIP(64)
    idx = faiss.IndexIVFFlat(quant, 64, 32, faiss.METRIC_INNER_PRODUCT)
    idx.train(x)
    idx.add(x)
    idx.nprobe = 4
    D, I = idx.search(q, 50)
    o = OracleIndexIVFFlat(64, 32)
    o.set_centroids(faiss.vector_to_array(quant.get_xb() if hasattr(quant, "get_xb") else quant.xb).reshape(32, 64))
    o.add(x)
    o.nprobe = 4
    rd, rid = o.search(q, 50, extra=16)
    compare_topk(I, D, rid, rd, 50, gap_tol=1e-6)
