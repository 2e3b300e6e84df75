"""GPU parity tests of the Flat path: every call goes FAISSIndex -> ctypes -> libb2retr.so
(C ABI) -> sm_100a kernels, and is compared with the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GAP_TOL = 1e-6      # adjacent-score gap below which order may differ from the oracle (fp32 accumulation order)
SCORE_RTOL = 1e-3   # north_star tolerance; after the fp32 rescore the observed error is ~1e-7


@pytest.fixture(scope="module")
def fr(built_lib):
    import torch
    assert torch.cuda.is_available()
    from movie_recommender_demo_b200 import faiss_retrieval
    faiss_retrieval.FAISSIndex.verbose = False
    return faiss_retrieval


def _bf16(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def _parity(fr, N, Q, k, d=256, seed=0, force_path=0, unit=False, epi_warps=0, walk=None, pair_scan=None):
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, d)).astype(np.float32)
    q = rng.standard_normal((Q, d)).astype(np.float32)
    g = fr.FAISSIndex(d, 'Flat')
    if force_path:
        g.index.set_param("force_path", force_path)
    if epi_warps:
        g.index.set_param("epi_warps", epi_warps)
    if walk is not None:
        g.index.set_param("walk", walk)
    if pair_scan is not None:
        g.index.set_param("pair_scan", pair_scan)
    g.add(x)
    o = OracleFAISSIndex(d, 'Flat')
    o.add(x)
    ids, dist = g.search(q, k=k)
    rid, rd = o.search(q, k=k, extra=32)
    res = compare_topk(ids, dist, rid, rd, k, gap_tol=GAP_TOL, score_rtol=SCORE_RTOL)
    assert (g.index.last_status == 0).all()
    assert dist.dtype == np.float32 and ids.dtype == np.int64
    return res


def _round16(a, fp16):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return t.to(torch.float16 if fp16 else torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("normalize", [False, True])      # un-normalised add -> bf16 scan copy, normalised -> fp16
@pytest.mark.parametrize("N,Q,d", [(256, 16, 256), (1000, 1, 256), (5000, 130, 256), (70001, 300, 128),
                                   (33, 5, 192), (100000, 64, 64)])
def test_tcgen05_score_tile_matches_16bit_reference(fr, N, Q, d, normalize):
    """Dump mode of the scan kernel == (unit-norm query, 16-bit) x (stored 16-bit rows), fp32 accumulate."""
    from oracle.flat import normalize_L2
    rng = np.random.default_rng(N + Q)
    x = rng.standard_normal((N, d)).astype(np.float32)
    q = rng.standard_normal((Q, d)).astype(np.float32)
    idx = fr.IndexFlatIP(d)
    idx.add(x, normalize=normalize)
    fp16 = int(idx.get_param("scan_dtype")) == 1
    assert fp16 == normalize
    tc = idx.debug_scores(q, "tc").cpu().numpy()
    simt = idx.debug_scores(q, "simt").cpu().numpy()
    xs = idx.reconstruct_n(0, N).cpu().numpy()          # the stored fp32 rows (the 16-bit copy rounds exactly these)
    ref = _round16(normalize_L2(q.copy()), fp16).astype(np.float64) @ _round16(xs, fp16).astype(np.float64).T
    scale = np.abs(ref).max()
    # fp32 accumulation-order noise + the odd 16-bit rounding flip of a query element: the GPU normalises
    # the query with its own summation order, and a 1-ulp fp32 difference can flip one 16-bit rounding
    # (<= 2^-11 |q_i||x_i| per score for fp16).  A layout / descriptor bug would be O(scale).
    tol = 2e-6 * scale * np.sqrt(d) + 1e-3 * scale
    assert np.abs(tc - ref).max() <= tol
    assert np.abs(simt - ref).max() <= tol
    assert np.median(np.abs(tc - ref)) <= 1e-6 * scale


@pytest.mark.parametrize("scan_dtype", [0, 1])
def test_both_scan_formats_give_the_exact_answer(fr, scan_dtype):
    """bf16 (north_star wording) and fp16 scan copies must both reproduce the fp32 oracle; fp16 needs a
    smaller rescore window."""
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex
    rng = np.random.default_rng(21)
    x = rng.standard_normal((300000, 256)).astype(np.float32)
    q = rng.standard_normal((16, 256)).astype(np.float32)
    g = fr.FAISSIndex(256, 'Flat')
    g.index.set_param("scan_dtype", scan_dtype)
    g.index.set_param("force_path", 2)
    g.add(x)
    assert int(g.index.get_param("scan_dtype")) == scan_dtype
    ids, dist = g.search(q, k=500)
    o = OracleFAISSIndex(256, 'Flat')
    o.add(x)
    rid, rd = o.search(q, k=500, extra=32)
    compare_topk(ids, dist, rid, rd, 500, gap_tol=GAP_TOL)
    assert (g.index.last_status == 0).all()


def test_unnormalised_faiss_level_index_switches_to_bf16_and_stays_exact(fr):
    """index.add(x) without normalisation (plain faiss IndexFlatIP semantics) after a normalised add
    re-encodes the scan copy as bf16; large-magnitude rows and un-normalised queries stay exact."""
    rng = np.random.default_rng(22)
    a = rng.standard_normal((3000, 64)).astype(np.float32)
    b = (rng.standard_normal((2000, 64)) * 300.0).astype(np.float32)      # would overflow an fp16 scan copy
    idx = fr.IndexFlatIP(64)
    idx.add(a, normalize=True)
    assert int(idx.get_param("scan_dtype")) == 1
    idx.add(b, normalize=False)
    assert int(idx.get_param("scan_dtype")) == 0
    q = (rng.standard_normal((5, 64)) * 7.0).astype(np.float32)
    D, I = idx.search(q, 40, normalize=False)
    xs = np.concatenate([a / np.linalg.norm(a, axis=1, keepdims=True), b])
    S = q @ xs.T
    ref = np.argsort(-S, axis=1, kind="stable")[:, :40]
    assert np.array_equal(I, ref)
    np.testing.assert_allclose(D, np.take_along_axis(S, ref, axis=1), rtol=2e-6)


@pytest.mark.parametrize("N,Q,k", [(20000, 37, 50), (3000, 5, 500), (100, 3, 500), (1, 2, 10), (4096, 8, 100)])
def test_dense_path_matches_oracle(fr, N, Q, k):
    _parity(fr, N, Q, k, seed=N, force_path=1)
    _parity(fr, N, Q, k, seed=N + 7)            # whichever path the planner picks for this size


@pytest.mark.parametrize("N,Q,k", [(45000, 9, 500), (60000, 130, 500), (100000, 1, 500), (8000, 4, 100),
                                   (20000, 700, 200), (250000, 33, 1000)])
def test_small_corpus_filter_path_thresholds_hold(fr, N, Q, k):
    """Corpora just above the dense cut-off: candidates are dense in the sampled 32-row groups (the
    sample rank is collision-corrected); the answer must match the oracle with no inexact flag, and the
    planner's threshold should rarely need the retry path."""
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex
    rng = np.random.default_rng(N + k)
    x = rng.standard_normal((N, 256)).astype(np.float32)
    q = rng.standard_normal((Q, 256)).astype(np.float32)
    g = fr.FAISSIndex(256, 'Flat')
    g.add(x)
    o = OracleFAISSIndex(256, 'Flat')
    o.add(x)
    ids, dist = g.search(q, k=k)
    rid, rd = o.search(q, k=k, extra=32)
    compare_topk(ids, dist, rid, rd, k, gap_tol=GAP_TOL, score_rtol=SCORE_RTOL)
    assert (g.index.last_status == 0).all()
    assert g.index.last_retries <= 1


def test_config1_shape_100k_ads_512_queries_top500(fr):
    """BASELINE config 1 shape: IndexFlatIP top-500 over 100k embeddings, 512 queries."""
    res = _parity(fr, 100000, 512, 500, seed=11)
    assert res["exact_positions"] > 0.95 * 512 * 500


@pytest.mark.parametrize("N,Q,k,force", [(300000, 8, 100, 2), (1000000, 64, 500, 0), (1000000, 1, 500, 0),
                                          (600000, 300, 500, 0), (150000, 3, 10, 2)])
def test_filter_path_matches_oracle(fr, N, Q, k, force):
    _parity(fr, N, Q, k, seed=N + Q, force_path=force)


@pytest.mark.parametrize("epi_warps", [8, 16])
@pytest.mark.parametrize("N,Q,k,d", [(400000, 300, 500, 256), (250000, 1000, 100, 128), (999999, 257, 500, 64)])
def test_two_query_block_filter_scan_both_epilogue_layouts(fr, N, Q, k, d, epi_warps):
    """Batches above 128 queries run the MQ = 2 filter scan; its 8-warp (one candidate segment per corpus
    split) and 16-warp (one per 64-column half) epilogues must both give the oracle's answer."""
    _parity(fr, N, Q, k, d=d, seed=N + Q, epi_warps=epi_warps)


@pytest.mark.parametrize("pair_scan", [0, 1])
@pytest.mark.parametrize("N,Q,k,d", [(400000, 300, 500, 256), (250000, 1000, 100, 128), (999999, 257, 500, 64),
                                     (300001, 129, 500, 192), (65000, 600, 500, 256)])
def test_cta_pair_filter_scan_matches_oracle(fr, N, Q, k, d, pair_scan):
    """`pair_scan = 1` (default) runs the filter scan of batches above 128 queries on CTA pairs (tcgen05 cta_group::2:
    256 queries x 256 corpus rows per UMMA, scan_pair.cu); 0 is the one-CTA kernel.  Both give the oracle's answer
    (odd corpus sizes: the last pair tile is half or partly empty; Q = 129 / 257: an all-padding second block)."""
    _parity(fr, N, Q, k, d=d, seed=N + Q, pair_scan=pair_scan)


def test_cta_pair_scan_and_one_cta_scan_return_identical_results(fr):
    import torch
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((700_003, 256), generator=g, device="cuda")
    q = torch.randn((1500, 256), generator=g, device="cuda")
    idx = fr.IndexFlatIP(256)
    idx.add(x, normalize=True)
    out = []
    for pair in (1, 0):
        idx.set_param("pair_scan", pair)
        assert int(idx.get_param("pair_scan")) == pair
        D, I, st, _ = idx.search_device(q, 500, normalize=True)
        assert int((st != 0).sum()) == 0
        out.append((D.clone(), I.clone()))
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][0], out[1][0])


@pytest.mark.parametrize("walk", [0, 1])
@pytest.mark.parametrize("N,Q,k", [(500000, 300, 500), (300000, 8, 100), (1000000, 128, 500)])
def test_filter_hit_walk_variants(fr, N, Q, k, walk):
    """The append walk of the filter epilogue (all 8 scores of a passing group, or only its passing
    3-element sub-groups) is an implementation detail: both must give the oracle's answer."""
    _parity(fr, N, Q, k, seed=N + Q + 1, walk=walk)


def test_ingest_normalises_like_faiss_and_never_mutates_input(fr):
    from oracle.flat import normalize_L2
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((1000, 256)) * 3).astype(np.float64)
    x[5] = 0
    x0 = x.copy()
    g = fr.FAISSIndex(256, 'Flat')
    g.add(x)
    assert (x == x0).all()
    got = g.index.reconstruct_n(0, 1000).cpu().numpy()
    ref = normalize_L2(x.astype(np.float32))
    assert np.abs(got - ref).max() < 2e-7
    assert (got[5] == 0).all()                      # zero-norm rows stay zero
    raw = fr.IndexFlatIP(256)
    raw.add(x.astype(np.float32))                   # faiss-level add: no normalisation, bit-exact copy
    assert np.array_equal(raw.reconstruct_n(0, 1000).cpu().numpy(), x.astype(np.float32))


def test_wrapper_semantics(fr):
    """ids first, default ids continue, k > ntotal pads with id_map[-1] / -FLT_MAX, object ids."""
    from oracle.flat import NEG_FLT_MAX, OracleFAISSIndex
    rng = np.random.default_rng(6)
    x = rng.standard_normal((300, 64)).astype(np.float32)
    q = rng.standard_normal((4, 64)).astype(np.float32)
    g = fr.FAISSIndex(64, 'Flat')
    o = OracleFAISSIndex(64, 'Flat')
    for t in (g, o):
        t.add(x[:100])
        t.add(x[100:], ad_ids=list(range(1000, 1200)))
    assert g.id_map == o.id_map and g.index.ntotal == 300 and g.index.is_trained
    ids, d = g.search(q, k=320)
    oid, od = o.search(q, k=320)
    assert np.array_equal(ids[:, :300], oid[:, :300])
    assert (ids[:, 300:] == 1199).all() and (d[:, 300:] == NEG_FLT_MAX).all()
    np.testing.assert_allclose(d[:, :300], od[:, :300], atol=3e-7)
    assert g.search(q, k=5, return_distances=False).shape == (4, 5)
    bi, bd = g.batch_search(q, k=7, batch_size=3)
    assert np.array_equal(bi, oid[:, :7])
    assert g.get_stats() == {'index_type': 'Flat', 'dimension': 64, 'num_vectors': 300, 'is_trained': True,
                             'nlist': 100, 'nprobe': 10}
    # arbitrary python ids -> host mapping, still id_map[idx]
    s = fr.FAISSIndex(64, 'Flat')
    s.add(x[:50], ad_ids=[f"ad{i}" for i in range(50)])
    sid, _ = s.search(q, k=3)
    so = OracleFAISSIndex(64, 'Flat')
    so.add(x[:50], ad_ids=[f"ad{i}" for i in range(50)])
    assert np.array_equal(sid, so.search(q, k=3)[0])
    # faiss-level object: (D, I) order; labels when no id map is attached
    raw = fr.IndexFlatIP(64)
    raw.add(x, normalize=True)
    D, I = raw.search(q, 5, normalize=True)
    assert D.shape == (4, 5) and I.dtype == np.int64 and I.max() < 300
    assert np.array_equal(np.asarray(g.id_map)[I], ids[:, :5])


def test_wrapper_matches_the_reference_wrapper_golden(fr):
    """The B200 `FAISSIndex` against tests/golden/wrapper_flat.npz = what the reference's OWN wrapper class returns
    (faiss_retrieval.py:14-256 executed unmodified over a numpy stand-in for faiss's IndexFlatIP / normalize_L2,
    tests/golden/make_wrapper_golden.py): float64 input, custom then default ids, k > ntotal, ids-only return,
    batch_search, get_stats; the caller's arrays are never written."""
    from oracle.compare import compare_topk
    fx = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "wrapper_flat.npz")
    x1, x2, q = fx["x1"], fx["x2"], fx["q"]
    x1c, x2c, qc = x1.copy(), x2.copy(), q.copy()
    g = fr.FAISSIndex(64, 'Flat')
    g.add(x1, ad_ids=fx["ids1"].tolist())
    g.add(x2)
    assert list(g.id_map) == fx["id_map"].tolist()
    ids, dist = g.search(q[:9], k=10)
    compare_topk(ids, dist, fx["ids_k10"], fx["dist_k10"], 10, gap_tol=1e-6)
    ids, dist = g.search(q[:3], k=520)                       # 20 slots past ntotal: id_map[-1] / -FLT_MAX
    compare_topk(ids, dist, fx["ids_big"], fx["dist_big"], 520, gap_tol=1e-6)
    assert (ids[:, 500:] == fx["id_map"][-1]).all() and np.array_equal(dist[:, 500:], fx["dist_big"][:, 500:])
    only = g.search(q[:4], k=6, return_distances=False)
    assert only.shape == (4, 6) and all(set(a) == set(b) for a, b in zip(only.tolist(), fx["ids_only"].tolist()))
    b_ids, b_dist = g.batch_search(q, k=5, batch_size=7)
    compare_topk(b_ids, b_dist, fx["batch_ids"], fx["batch_dist"], 5, gap_tol=1e-6)
    stats = g.get_stats()
    assert list(stats.keys()) == fx["stats_keys"].tolist()
    assert [str(v) for v in stats.values()] == fx["stats_values"].tolist()
    assert np.array_equal(x1, x1c) and np.array_equal(x2, x2c) and np.array_equal(q, qc)


def test_accepts_cuda_tensors_and_empty_inputs(fr):
    import torch
    g = fr.FAISSIndex(128, 'Flat')
    D, I = g.index.search(np.zeros((2, 128), np.float32), 4)
    assert (I == -1).all()                          # empty index
    x = torch.randn(5000, 128, device="cuda")
    g.add(x)
    q = torch.randn(3, 128, device="cuda")
    ids, d = g.search(q, k=10)
    ids2, d2 = g.search(q.cpu().numpy(), k=10)
    assert np.array_equal(ids, ids2) and np.array_equal(d, d2)
    ids0, d0 = g.index.search(np.zeros((0, 128), np.float32), 4)
    assert ids0.shape == (0, 4)


def test_duplicates_and_zero_queries(fr):
    """Exact duplicate rows tie -> canonical order by label; an all-zero query scores 0 everywhere."""
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex
    rng = np.random.default_rng(8)
    x = rng.standard_normal((2000, 64)).astype(np.float32)
    x[500:520] = x[7]
    q = rng.standard_normal((3, 64)).astype(np.float32)
    q[1] = x[7]
    g = fr.FAISSIndex(64, 'Flat')
    o = OracleFAISSIndex(64, 'Flat')
    g.add(x), o.add(x)
    ids, d = g.search(q, k=30)
    rid, rd = o.search(q, k=30, extra=32)
    compare_topk(ids, d, rid, rd, 30, gap_tol=GAP_TOL)
    assert set(ids[1, :21].tolist()) == {7, *range(500, 520)}


def test_full_size_round_trip_properties(fr):
    """BASELINE config 2 size (1M x 256, top-500): size-independent properties — every query that
    IS a corpus row retrieves itself first with score 1, scores are sorted, ids unique, and a
    second search returns bit-identical results."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(1)
    N, d = 1_000_000, 256
    x = torch.randn((N, d), generator=g, device="cuda")
    idx = fr.FAISSIndex(d, 'Flat')
    idx.add(x)
    rows = torch.tensor([0, 17, 123456, 999999], device="cuda")
    ids, dist = idx.search(x[rows], k=500)
    assert ids[:, 0].tolist() == rows.tolist()
    np.testing.assert_allclose(dist[:, 0], 1.0, atol=2e-6)
    assert (np.diff(dist, axis=1) <= 0).all()
    assert all(len(set(r.tolist())) == 500 for r in ids)
    ids2, dist2 = idx.search(x[rows], k=500)
    assert np.array_equal(ids, ids2) and np.array_equal(dist, dist2)
    # rescored scores equal an fp32 torch dot of the normalised rows
    xn = torch.nn.functional.normalize(x[rows], dim=1)
    ref = (xn[:, None, :] * torch.nn.functional.normalize(x[torch.from_numpy(ids).cuda()], dim=2)).sum(-1)
    np.testing.assert_allclose(dist, ref.cpu().numpy(), atol=3e-6)
    assert (idx.index.last_status == 0).all()


def test_threshold_retry_path_recovers(fr):
    """Force a far-too-high candidate threshold: the status bits must flag it and the host retry
    loop must still return the oracle's answer."""
    import torch
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex
    rng = np.random.default_rng(9)
    x = rng.standard_normal((200000, 64)).astype(np.float32)
    q = rng.standard_normal((4, 64)).astype(np.float32)
    g = fr.FAISSIndex(64, 'Flat')
    g.add(x)
    tau = torch.full((4,), 0.9, device="cuda")
    D, I, st, tr = g.index.search_device(q, 50, normalize=True, tau=tau)
    assert (st.cpu().numpy() & 1).all()             # B2R_ST_TOO_FEW
    g.index.set_param("cand_factor", 1.0)           # tight threshold -> NEED_LOWER_TAU retries likely
    g.index.set_param("force_path", 2)
    ids, d = g.search(q, k=50)
    o = OracleFAISSIndex(64, 'Flat')
    o.add(x)
    rid, rd = o.search(q, k=50, extra=32)
    compare_topk(ids, d, rid, rd, 50, gap_tol=GAP_TOL)
    assert (g.index.last_status == 0).all()


def test_topk_merge_matches_unsharded(fr, built_lib):
    """P logical shards on one device + b2r_topk_merge == unsharded search."""
    import torch
    from movie_recommender_demo_b200 import _lib
    rng = np.random.default_rng(10)
    N, d, Q, k, P = 40000, 64, 9, 100, 4
    x = rng.standard_normal((N, d)).astype(np.float32)
    q = rng.standard_normal((Q, d)).astype(np.float32)
    full = fr.IndexFlatIP(d)
    full.add(x, normalize=True)
    Dref, Iref = full.search(q, k, normalize=True)
    Ds, Is = [], []
    for s in range(P):
        lo, hi = s * N // P, (s + 1) * N // P
        sh = fr.IndexFlatIP(d)
        sh.add(x[lo:hi], normalize=True)
        sh.set_label_base(lo)
        D, I = sh.search(q, k, normalize=True, return_device=True)
        Ds.append(D), Is.append(I)
    D_all = torch.stack(Ds).contiguous()
    I_all = torch.stack(Is).contiguous()
    D_out = torch.empty((Q, k), dtype=torch.float32, device="cuda")
    I_out = torch.empty((Q, k), dtype=torch.int64, device="cuda")
    _lib.check(built_lib.b2r_topk_merge(P, Q, k, D_all.data_ptr(), I_all.data_ptr(), D_out.data_ptr(),
                                        I_out.data_ptr(), 1, int(torch.cuda.current_stream().cuda_stream)))
    assert np.array_equal(I_out.cpu().numpy(), Iref)
    assert np.array_equal(D_out.cpu().numpy(), Dref)


@pytest.mark.parametrize("d", [20, 32, 100, 130, 200, 255])
def test_any_dimension_up_to_256(fr, d, tmp_path):
    """The reference takes any `dimension` (faiss_retrieval.py:20-24); dimensions that are not a multiple
    of 64 are zero-padded on the way in and must give the oracle's answer, the stored rows back, and a
    faiss-layout file of the TRUE dimension."""
    from movie_recommender_demo_b200 import faiss_io
    from oracle.flat import normalize_L2
    _parity(fr, 30000, 19, 100, d=d, seed=d)
    _parity(fr, 400000, 6, 500, d=d, seed=d + 1, force_path=2)
    rng = np.random.default_rng(d)
    x = rng.standard_normal((500, d)).astype(np.float32)
    g = fr.FAISSIndex(d, 'Flat')
    g.add(x)
    back = g.index.reconstruct_n(0, 500).cpu().numpy()
    assert back.shape == (500, d) and np.allclose(back, normalize_L2(x.copy()), atol=1e-6)
    path = str(tmp_path / "odd.index")
    g.save(path)
    desc = faiss_io.parse(open(path, "rb"))
    assert desc["d"] == d and desc["xb"].shape == (500, d)
    h = fr.FAISSIndex(d, 'Flat')
    h.load(path)
    a, da = g.search(x[:5], k=20)
    b, db = h.search(x[:5], k=20)
    assert np.array_equal(a, b) and np.array_equal(da, db)
    with pytest.raises(ValueError, match="expected shape"):
        g.search(np.zeros((2, d + 1), np.float32), k=5)


def test_dimension_limits(fr):
    with pytest.raises(ValueError, match="not supported"):
        fr.FAISSIndex(257, 'Flat')
    with pytest.raises(ValueError, match="multiple of 64"):
        fr.FAISSIndex(100, 'IVFPQ', nlist=4)


def test_graph_replay_path_matches_eager(fr, monkeypatch):
    """Host-result searches of <= 256 queries replay a captured CUDA graph; the answers must be identical to
    the eager launch sequence, survive new query values / repeated calls, and be re-captured after the corpus
    changes (add), for every index family."""
    rng = np.random.default_rng(5)
    d = 128
    x = rng.standard_normal((90000, d)).astype(np.float32)
    for kind in ("Flat", "IVF", "IVFPQ"):
        g = fr.FAISSIndex(d, kind, nlist=32, nprobe=6, pq_m=16)
        g.add(x[:60000])
        for trial in range(3):
            q = rng.standard_normal((7, d)).astype(np.float32)
            ids_g, dist_g = g.search(q, k=100)
            assert any(e for e in g.index._graphs.values()), "graph path was not taken"
            monkeypatch.setenv("B2R_NO_GRAPHS", "1")
            ids_e, dist_e = g.search(q, k=100)
            monkeypatch.delenv("B2R_NO_GRAPHS")
            assert np.array_equal(ids_g, ids_e) and np.array_equal(dist_g, dist_e), (kind, trial)
        g.add(x[60000:])                       # corpus changed: stale graphs must go
        assert not g.index._graphs
        q = x[60000:60005] + 0.01 * rng.standard_normal((5, d)).astype(np.float32)
        ids, _ = g.search(q, k=10)
        if kind != "IVFPQ":
            assert ids[:, 0].tolist() == list(range(60000, 60005))
        ids_t, _ = g.search(__import__("torch").from_numpy(q).cuda(), k=10)     # CUDA tensor input, same graph
        assert np.array_equal(ids, ids_t)


def test_reserve_keeps_rows_and_avoids_regrowth(fr):
    import torch
    rng = np.random.default_rng(9)
    x = rng.standard_normal((5000, 64)).astype(np.float32)
    g = fr.FAISSIndex(64, 'Flat')
    g.add(x[:1000])
    g.index.reserve(5000)
    before = torch.cuda.memory_allocated()
    g.add(x[1000:])
    assert g.index.ntotal == 5000
    ids, _ = g.search(x[[3, 4999]], k=1)
    assert ids[:, 0].tolist() == [3, 4999]
    g.index.reserve(10)            # smaller than the current size: no-op
    assert g.index.ntotal == 5000
    with pytest.raises(Exception):
        g.index.reserve(-1)
    assert before >= 0


def test_benchmark_faiss_index_surface(fr):
    """faiss_retrieval.benchmark_faiss_index (reference :372-437): same call, same result keys for the
    four index families."""
    res = fr.benchmark_faiss_index(dimension=64, num_vectors=20000, num_queries=7, k=10)
    assert set(res) == {'Flat', 'IVF', 'IVFPQ', 'HNSW'}
    for m in res.values():
        assert set(m) == {'add_time', 'search_time_ms', 'per_query_ms'} and m['search_time_ms'] > 0


@pytest.mark.parametrize("N,Q,k,d", [(30000, 9, 100, 256), (120000, 40, 500, 128), (300, 3, 500, 100)])
def test_hnsw_type_is_an_exact_l2_search(fr, N, Q, k, d, tmp_path):
    """index_type='HNSW' (faiss_retrieval.py:65-70: IndexHNSWFlat, L2, ascending): served by the exact scan,
    so it must equal the exact L2 oracle (= what HNSW approximates; recall 1.0), ids first, ascending
    distances, +FLT_MAX / id_map[-1] in unfilled slots, and survive save -> load."""
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex
    rng = np.random.default_rng(N)
    x = rng.standard_normal((N, d)).astype(np.float32)
    q = rng.standard_normal((Q, d)).astype(np.float32)
    g = fr.FAISSIndex(d, 'HNSW')
    assert (g.index.hnsw.M, g.index.hnsw.efConstruction, g.index.hnsw.efSearch) == (32, 40, 16)
    assert g.index.is_trained and not hasattr(g.index, "nprobe")
    ids0 = [7 * i + 1 for i in range(N)]
    g.add(x, ids0)
    o = OracleFAISSIndex(d, 'HNSW')
    o.add(x, ids0)
    ids, dist = g.search(q, k=k)
    rid, rd = o.search(q, k=k, extra=32 if N > k + 32 else 0)
    assert (np.diff(dist[:, :min(k, N)], axis=1) >= 0).all()
    compare_topk(ids, dist, rid, rd, k, gap_tol=4e-6, score_rtol=1e-4, score_atol=4e-6, descending=False)
    if N < k:                                   # under-filled: faiss gives (-1, +FLT_MAX); wrapper maps -1 -> id_map[-1]
        assert np.array_equal(ids[:, :N], rid[:, :N])
        assert (ids[:, N:] == ids0[-1]).all() and (dist[:, N:] == np.float32(3.4028234663852886e38)).all()
    p = str(tmp_path / "hnsw.index")
    g.save(p)                                   # native container (no graph to put in an IHNf file)
    g2 = fr.FAISSIndex(d, 'Flat')
    g2.load(p)
    assert g2.index_type == 'HNSW' and type(g2.index).__name__ == 'IndexHNSWFlat' and g2.index.ntotal == N
    ids2, dist2 = g2.search(q, k=k)
    assert np.array_equal(ids, ids2) and np.array_equal(dist, dist2)


def test_hnsw_rejects_rows_it_cannot_rank_and_reads_faiss_files(fr, tmp_path):
    import pickle
    from test_faiss_io_cpu import _hnsw_file
    rng = np.random.default_rng(3)
    h = fr.IndexHNSWFlat(64)
    with pytest.raises(ValueError, match="unit-norm"):
        h.add(rng.standard_normal((10, 64)).astype(np.float32))              # raw add of un-normalised rows
    with pytest.raises(ValueError, match="unit-norm"):
        h.add(np.zeros((2, 64), np.float32), normalize=True)                 # zero rows cannot be normalised
    assert h.ntotal == 0
    # an IndexHNSWFlat file authored from the published layout: vectors are taken, the graph is skipped
    x = rng.standard_normal((5000, 64)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    p = tmp_path / "ref_hnsw.index"
    p.write_bytes(_hnsw_file(x, M=32))
    with open(str(p) + ".metadata", "wb") as f:
        pickle.dump({'dimension': 64, 'index_type': 'HNSW', 'nlist': 100, 'nprobe': 10,
                     'id_map': list(range(5000))}, f)
    g = fr.FAISSIndex(64, 'Flat')
    g.load(str(p))
    assert g.index_type == 'HNSW' and g.index.ntotal == 5000 and g.index.hnsw.M == 32
    ids, dist = g.search(x[[11, 4321]], k=5)
    assert ids[:, 0].tolist() == [11, 4321] and np.allclose(dist[:, 0], 0.0, atol=2e-6)
    want = np.sort(((x[11][None] - x) ** 2).sum(1))[:5]
    assert np.allclose(dist[0], want, atol=4e-6)


@pytest.mark.parametrize("P,Q,k,q_rows", [(4, 9, 100, 9), (8, 37, 500, 40), (2, 300, 64, 300), (3, 5, 1000, 6)])
def test_packed_exchange_kernels_match_unsharded(fr, built_lib, P, Q, k, q_rows):
    """b2r_topk_pack + b2r_topk_merge_packed (the one-buffer, 8-bytes-per-result shard exchange, ranking merge)
    on P logical shards of one device == the unsharded search == b2r_topk_merge on the unpacked lists;
    status words are OR-ed; padding rows (q..q_rows) are empty lists."""
    import torch
    from movie_recommender_demo_b200 import _lib
    rng = np.random.default_rng(P * 1000 + Q)
    N, d = 50000, 64
    x = rng.standard_normal((N, d)).astype(np.float32)
    x[20000:20020] = x[3]                                   # exact ties, some across shard boundaries
    x[N // P - 3: N // P + 3] = x[5]
    q = rng.standard_normal((Q, d)).astype(np.float32)
    q[0] = x[3]
    q[1] = x[5]
    full = fr.IndexFlatIP(d)
    full.add(x, normalize=True)
    Dref, Iref = full.search(q, k, normalize=True)
    W = 2 * k + 1
    sp = int(torch.cuda.current_stream().cuda_stream)
    packed = torch.empty((P, q_rows, W), dtype=torch.int32, device="cuda")
    bases, Ds, Is, want_st = [], [], [], np.zeros(Q, dtype=np.int32)
    for s in range(P):
        lo, hi = s * N // P, (s + 1) * N // P
        sh = fr.IndexFlatIP(d)
        sh.add(x[lo:hi], normalize=True)
        sh.set_label_base(lo)
        D, I = sh.search(q, k, normalize=True, return_device=True)
        st = torch.zeros(Q, dtype=torch.int32, device="cuda")
        st[s::3] = 1 << (s % 4)
        want_st[s::3] |= 1 << (s % 4)
        _lib.check(built_lib.b2r_topk_pack(Q, q_rows, k, D.data_ptr(), I.data_ptr(), st.data_ptr(), lo,
                                           packed[s].data_ptr(), 1, sp))
        bases.append(lo), Ds.append(D), Is.append(I)
    bases_t = torch.tensor(bases, dtype=torch.int64, device="cuda")
    D_out = torch.empty((q_rows, k), dtype=torch.float32, device="cuda")
    I_out = torch.empty((q_rows, k), dtype=torch.int64, device="cuda")
    st_out = torch.empty(q_rows, dtype=torch.int32, device="cuda")
    _lib.check(built_lib.b2r_topk_merge_packed(P, q_rows, q_rows, k, packed.data_ptr(), bases_t.data_ptr(),
                                               D_out.data_ptr(), I_out.data_ptr(), st_out.data_ptr(), 1, sp))
    assert np.array_equal(I_out[:Q].cpu().numpy(), Iref)
    assert np.array_equal(D_out[:Q].cpu().numpy(), Dref)
    assert np.array_equal(st_out[:Q].cpu().numpy(), want_st)
    assert (I_out[Q:] == -1).all() and (st_out[Q:] == 0).all()
    if P * k <= 8192:
        D2 = torch.empty((Q, k), dtype=torch.float32, device="cuda")
        I2 = torch.empty((Q, k), dtype=torch.int64, device="cuda")
        D_all, I_all = torch.stack(Ds).contiguous(), torch.stack(Is).contiguous()      # kept alive across the launch
        _lib.check(built_lib.b2r_topk_merge(P, Q, k, D_all.data_ptr(), I_all.data_ptr(), D2.data_ptr(), I2.data_ptr(), 1, sp))
        torch.cuda.synchronize()
        assert torch.equal(I2, I_out[:Q]) and torch.equal(D2, D_out[:Q])
