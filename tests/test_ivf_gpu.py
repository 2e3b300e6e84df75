"""GPU parity of the IVF-Flat path (faiss_retrieval.py:50-55) against the oracle on SHARED
centroids: identical inverted-list membership, identical results, equal recall@k."""
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


@pytest.mark.parametrize("N,d,nlist,nprobe,Q,k", [(20000, 64, 32, 4, 17, 50), (60000, 256, 100, 10, 64, 500),
                                                   (5000, 128, 16, 16, 5, 100), (3000, 64, 8, 2, 300, 10)])
def test_ivf_matches_oracle_on_shared_centroids(fr, N, d, nlist, nprobe, Q, k):
    from oracle.compare import compare_topk, recall_at_k
    from oracle.flat import OracleFAISSIndex
    x = _clustered(N, d, nlist * 2, seed=N)
    q = _clustered(Q, d, nlist * 2, seed=N + 1)
    g = fr.FAISSIndex(d, 'IVF', nlist=nlist, nprobe=nprobe)
    assert not g.index.is_trained
    g.add(x)                                         # trains on the raw input first (faiss_retrieval.py:107-108)
    assert g.index.is_trained and g.index.ntotal == N
    o = OracleFAISSIndex(d, 'IVF', nlist=nlist, nprobe=nprobe)
    o.index.set_centroids(g.index.export_centroids())
    o.add(x)
    assert np.array_equal(g.index.list_sizes(), o.index.list_sizes()), "inverted-list membership differs"
    ids, dist = g.search(q, k=k)
    rid, rd = o.search(q, k=k, extra=32)
    compare_topk(ids, dist, rid, rd, k, gap_tol=1e-6)
    assert (g.index.last_status == 0).all()
    # recall@k against exact flat search is the same number for both (same ids)
    flat = OracleFAISSIndex(d, 'Flat')
    flat.add(x)
    truth, _ = flat.search(q, k=k)
    assert recall_at_k(ids, truth) == pytest.approx(recall_at_k(rid[:, :k], truth), abs=1e-3)


def test_ivf_default_reference_config_and_incremental_add(fr):
    """Reference defaults (nlist=100, nprobe=10, training_pipeline.py:531-536), two add() calls,
    under-filled probes return id_map[-1] / -FLT_MAX like the reference wrapper."""
    from oracle.compare import compare_topk
    from oracle.flat import NEG_FLT_MAX, OracleFAISSIndex
    d = 256
    x = _clustered(30000, d, 150, seed=3)
    q = _clustered(9, d, 150, seed=4)
    g = fr.FAISSIndex(d)                              # 'IVF', nlist=100, nprobe=10
    assert g.index_type == 'IVF' and g.nlist == 100 and g.nprobe == 10
    g.train(x)
    g.add(x[:20000])
    g.add(x[20000:], ad_ids=list(range(100000, 110000)))
    o = OracleFAISSIndex(d)
    o.index.set_centroids(g.index.export_centroids())
    o.add(x[:20000])
    o.add(x[20000:], ad_ids=list(range(100000, 110000)))
    assert g.id_map == o.id_map
    assert np.array_equal(g.index.list_sizes(), o.index.list_sizes())
    ids, dist = g.search(q, k=500)
    rid, rd = o.search(q, k=500, extra=32)
    compare_topk(ids, dist, rid, rd, 500, gap_tol=1e-6)
    # nprobe=1 with k larger than the probed list: unfilled slots
    g.nprobe = o.nprobe = 1
    ids1, d1 = g.search(q, k=1000)
    rid1, rd1 = o.search(q, k=1000)
    filled = (rd1 != NEG_FLT_MAX)
    assert (filled.sum(axis=1) < 1000).any()
    assert np.array_equal((d1 != NEG_FLT_MAX), filled)
    assert np.array_equal(ids1[~filled], rid1[~filled])          # id_map[-1]
    assert g.get_stats()['num_vectors'] == 30000 and g.get_stats()['is_trained']


def test_ivf_kmeans_quality_and_save_load(fr, tmp_path):
    """The GPU k-means must produce a usable quantiser (balanced-ish lists, high recall on clustered
    data) and save/load must round-trip the trained index."""
    from oracle.compare import recall_at_k
    from oracle.flat import OracleFAISSIndex
    d, N = 64, 40000
    x = _clustered(N, d, 64, seed=7)
    q = _clustered(50, d, 64, seed=8)
    g = fr.FAISSIndex(d, 'IVF', nlist=64, nprobe=8)
    g.add(x)
    sizes = g.index.list_sizes()
    assert sizes.sum() == N and (sizes > 0).mean() > 0.9
    flat = OracleFAISSIndex(d, 'Flat')
    flat.add(x)
    truth, _ = flat.search(q, k=100)
    ids, dist = g.search(q, k=100)
    assert recall_at_k(ids, truth) > 0.9
    path = str(tmp_path / "ivf.bin")
    g.save(path)
    h = fr.FAISSIndex(d, 'Flat')
    h.load(path)
    assert h.index_type == 'IVF' and h.index.ntotal == N and h.nprobe == 8
    ids2, dist2 = h.search(q, k=100)
    assert np.array_equal(ids, ids2) and np.array_equal(dist, dist2)


def test_ivf_odd_dimension(fr):
    """IVF-Flat with a dimension that is not a multiple of 64 (zero-padded internally): centroids come
    back at the true dimension and the answer matches the oracle on shared centroids."""
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex
    d, N, nlist = 100, 30000, 32
    x = _clustered(N, d, 64, seed=21)
    q = _clustered(23, d, 64, seed=22)
    g = fr.FAISSIndex(d, 'IVF', nlist=nlist, nprobe=6)
    g.add(x)
    cent = g.index.export_centroids()
    assert cent.shape == (nlist, d)
    o = OracleFAISSIndex(d, 'IVF', nlist=nlist, nprobe=6)
    o.index.set_centroids(cent)
    o.add(x)
    assert np.array_equal(g.index.list_sizes(), o.index.list_sizes())
    ids, dist = g.search(q, k=200)
    rid, rd = o.search(q, k=200, extra=32)
    compare_topk(ids, dist, rid, rd, 200, gap_tol=1e-6)


@pytest.mark.parametrize("kind,k", [("IVF", 500), ("IVF", 100), ("IVFPQ", 500)])
def test_sampled_threshold_never_changes_the_answer(fr, kind, k):
    """The IVF candidate threshold estimated from a score sample (default) must return exactly what the
    exact radix passes return: same ids, same distances, no status flags."""
    d, N, nlist = 128, 200000, 64
    x = _clustered(N, d, 128, seed=31)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = _clustered(150, d, 128, seed=32)
    g = fr.FAISSIndex(d, kind, nlist=nlist, nprobe=16, pq_m=16)
    g.add(x)
    out = {}
    for sample in (1, 0):
        g.index.set_param("ivf_sample", sample)
        assert g.index.get_param("ivf_sample") == sample
        out[sample] = g.search(q, k=k)
        assert (g.index.last_status == 0).all()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("N,d,nlist,nprobe,Q,k", [(200000, 256, 256, 16, 1, 500), (200000, 256, 256, 16, 64, 500),
                                                   (200000, 256, 256, 16, 1100, 500), (50000, 128, 64, 8, 1033, 100),
                                                   (300000, 64, 128, 128, 1040, 10), (9000, 256, 20, 3, 1300, 500),
                                                   (400000, 256, 512, 32, 4096, 500)])
def test_fused_list_scan_returns_what_the_dump_path_returns(fr, N, d, nlist, nprobe, Q, k):
    """IVF-Flat default: sample pass -> per-query threshold -> threshold filter inside the list scan's epilogue
    (no pair score reaches HBM; taken for chunks of 1024 queries and more).  `ivf_fused = 0` dumps every pair score
    and selects afterwards.  Same corpus, same
    centroids: ids and scores must be identical, with no query left flagged (flagged ones are re-run through the
    dump path by the wrapper)."""
    import torch
    from movie_recommender_demo_b200 import ivf
    x = torch.from_numpy(_clustered(N, d, nlist * 2, seed=N + d))
    q = torch.from_numpy(_clustered(Q, d, nlist * 2, seed=N + d + 1))
    idx = ivf.IndexIVFFlat(d, nlist)
    idx.train(x[: min(N, 64 * nlist)].cuda())
    idx.add(x.cuda(), normalize=True)
    out = []
    for fused in (1, 0):
        idx.set_param("ivf_fused", fused)
        assert int(idx.get_param("ivf_fused")) == fused
        D, I = idx.search(q.numpy(), k, normalize=True, nprobe=nprobe)
        assert (idx.last_status == 0).all()
        out.append((D, I, idx.last_retries))
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][0], out[1][0])
    assert out[1][2] == 0


def test_fused_list_scan_falls_back_when_the_threshold_misses(fr):
    """k = 1000 against a candidate target of 2560: the sampled threshold can let fewer candidates through than
    the top-k and its rescore window need for some queries; those are flagged and re-run through the dump path -
    the answer is still the exact one."""
    import torch
    from movie_recommender_demo_b200 import ivf
    N, d, nlist, nprobe, Q, k = 150000, 128, 64, 16, 1200, 1000
    x = torch.from_numpy(_clustered(N, d, 200, seed=77)).cuda()
    q = _clustered(Q, d, 200, seed=78)
    idx = ivf.IndexIVFFlat(d, nlist)
    idx.train(x[:8192])
    idx.add(x, normalize=True)
    D1, I1 = idx.search(q, k, normalize=True, nprobe=nprobe)
    retried, st1 = idx.last_retries, idx.last_status.copy()
    idx.set_param("ivf_fused", 0)
    D0, I0 = idx.search(q, k, normalize=True, nprobe=nprobe)
    assert (st1 == 0).all() and (idx.last_status == 0).all()
    assert np.array_equal(I1, I0) and np.array_equal(D1, D0)
    assert retried in (0, 1)
