"""CPU tests: the oracle against the golden fixtures / brute force, and the comparator."""
import numpy as np
import pytest

from oracle import towers as otowers
from oracle.compare import compare_topk, recall_at_k
from oracle.flat import NEG_FLT_MAX, OracleFAISSIndex, OracleIndexFlatIP, normalize_L2, topk_desc
from weights import CONFIGS, DEPTH_CONFIGS, make_inputs, make_state

GOLDEN = __import__("pathlib").Path(__file__).parent / "golden"


@pytest.mark.parametrize("cfg_name", ["cfg1", "small"] + DEPTH_CONFIGS)
def test_oracle_towers_match_reference_golden(cfg_name):
    """oracle/towers.py vs outputs of the reference's own two_tower_model.py."""
    fx = np.load(GOLDEN / f"towers_{cfg_name}.npz")
    cfg = CONFIGS[cfg_name]
    state = make_state(cfg, int(fx["seed"]))
    assert sorted(state.keys()) == sorted(fx["keys"].tolist())
    ucat, unum, acat = make_inputs(cfg, int(fx["seed"]), fx["ucat"].shape[0])
    assert (ucat == fx["ucat"]).all() and (acat == fx["acat"]).all() and (unum == fx["unum"]).all()
    # embedding gather + concat: bit exact
    assert np.array_equal(otowers.embedding_concat(state, "user_tower", ucat), fx["user_embedding_layer"])
    assert np.array_equal(otowers.embedding_concat(state, "ad_tower", acat), fx["ad_embedding_layer"])
    u = otowers.tower_forward(state, "user_tower", ucat, unum)
    a = otowers.tower_forward(state, "ad_tower", acat)
    np.testing.assert_allclose(u, fx["user_out"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(a, fx["ad_out"], rtol=0, atol=5e-6)


def test_embedding_out_of_range_raises():
    cfg = CONFIGS["small"]
    state = make_state(cfg, 1)
    ucat, _, _ = make_inputs(cfg, 1, 4)
    ucat[2, 1] = cfg["user_cards"][1]
    with pytest.raises(IndexError):
        otowers.embedding_concat(state, "user_tower", ucat)


def test_normalize_l2_semantics():
    x = np.array([[3, 4], [0, 0], [1e-30, 0]], dtype=np.float32)
    y = normalize_L2(x.copy())
    np.testing.assert_allclose(y[0], [0.6, 0.8], rtol=1e-6)
    assert (y[1] == 0).all()          # zero rows stay zero (faiss fvec_renorm_L2)


def test_flat_topk_matches_bruteforce_and_canonical_ties():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((5000, 32)).astype(np.float32)
    X[100] = X[7]                       # exact duplicate -> tie broken by label
    Q = rng.standard_normal((9, 32)).astype(np.float32)
    o = OracleFAISSIndex(32, 'Flat')
    o.add(X)
    ids, d = o.search(Q, k=40)
    Xn = normalize_L2(X.copy())
    Qn = normalize_L2(Q.copy())
    S = Qn @ Xn.T
    ref = np.argsort(-S, axis=1, kind="stable")[:, :40]
    assert np.array_equal(ids, ref)
    assert np.array_equal(d, np.take_along_axis(S, ref, axis=1))
    assert (np.diff(d, axis=1) <= 0).all()


def test_wrapper_semantics_ids_and_empty_slots():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((30, 8)).astype(np.float64)   # fp64 accepted, never mutated
    X0 = X.copy()
    o = OracleFAISSIndex(8, 'Flat')
    o.add(X[:10])
    o.add(X[10:], ad_ids=[f"ad{i}" for i in range(20)])   # arbitrary python ids
    assert (X == X0).all()
    assert o.id_map[:10] == list(range(10)) and o.id_map[10] == "ad0"   # default ids continue from len(id_map)
    ids, d = o.search(rng.standard_normal((2, 8)), k=40)
    assert ids.shape == (2, 40)
    assert (ids[:, 30:] == "ad19").all()                  # label -1 -> id_map[-1] (faiss_retrieval.py:159-160)
    assert (d[:, 30:] == NEG_FLT_MAX).all()
    only_ids = o.search(rng.standard_normal((2, 8)), k=5, return_distances=False)
    assert only_ids.shape == (2, 5)
    ids_b, d_b = o.batch_search(rng.standard_normal((7, 8)), k=3, batch_size=2)
    assert ids_b.shape == (7, 3)
    with pytest.raises(ValueError, match="Unknown index type: Bogus"):
        OracleFAISSIndex(8, 'Bogus')
    assert o.get_stats()["num_vectors"] == 30


def test_comparator_accepts_near_tie_permutations_only():
    ref_d = np.array([[0.9, 0.5, 0.5 - 5e-7, 0.3, 0.2, 0.2 - 1e-7]], dtype=np.float32)
    ref_i = np.array([[4, 8, 2, 7, 1, 9]])
    k = 5
    ok_i = np.array([[4, 2, 8, 7, 9]])       # swap inside the tie run; boundary run may pick 9 instead of 1
    ok_d = ref_d[:, :k]
    r = compare_topk(ok_i, ok_d, ref_i, ref_d, k)
    assert r["tie_positions"] == 3 and r["exact_positions"] == 2
    with pytest.raises(AssertionError):
        compare_topk(np.array([[8, 4, 2, 7, 1]]), ok_d, ref_i, ref_d, k)     # real order violation
    with pytest.raises(AssertionError):
        compare_topk(np.array([[4, 8, 2, 7, 3]]), ok_d, ref_i, ref_d, k)     # foreign id
    bad_d = ok_d.copy()
    bad_d[0, 0] = 0.95
    with pytest.raises(AssertionError):
        compare_topk(ok_i, bad_d, ref_i, ref_d, k)


def test_topk_desc_k_larger_than_n():
    S = np.array([[0.1, 0.9, 0.5]], dtype=np.float32)
    D, I = topk_desc(S, 5)
    assert I.tolist() == [[1, 2, 0, -1, -1]] and (D[0, 3:] == NEG_FLT_MAX).all()
    assert recall_at_k(I, I) == 1.0


# ---- self-consistency of the IVF / IVF-PQ restatements (they are unpinned against faiss: at least pin them
# ---- against the definitions they claim to restate)
def _unit(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def test_oracle_ivf_probing_every_list_is_the_flat_answer():
    from oracle.flat import OracleIndexFlatIP
    from oracle.ivf import OracleIndexIVFFlat
    x, q = _unit(3000, 32, 1), _unit(9, 32, 2)
    ivf = OracleIndexIVFFlat(32, 16)
    ivf.train(x)
    ivf.add(x)
    assert ivf.list_sizes().sum() == 3000
    flat = OracleIndexFlatIP(32)
    flat.add(x)
    ivf.nprobe = 16
    D1, I1 = ivf.search(q, 40)
    D2, I2 = flat.search(q, 40)
    assert np.array_equal(I1, I2) and np.allclose(D1, D2, atol=1e-6)
    # fewer probes: every returned row lies in a probed list, scores are exact, recall can only go down
    ivf.nprobe = 3
    D3, I3 = ivf.search(q, 40)
    probed = ivf.probe(q)
    for qi in range(len(q)):
        ok = I3[qi] >= 0
        assert np.isin(ivf.assign[I3[qi][ok]], probed[qi]).all()
        assert np.allclose(D3[qi][ok], x[I3[qi][ok]] @ q[qi], atol=1e-6)
        assert len(set(I3[qi][ok]) & set(I2[qi])) <= 40


def test_oracle_ivfpq_adc_is_the_distance_to_the_decoded_vector():
    """ADC distance = |q - (centroid + decoded residual)|^2 (by-residual, L2), ascending, canonical ties."""
    from oracle.ivfpq import OracleIndexIVFPQ
    x, q = _unit(4000, 32, 3), _unit(6, 32, 4)
    pq = OracleIndexIVFPQ(32, 8, m=4)
    pq.train(x)
    pq.add(x)
    assert pq.codes.shape == (4000, 4) and pq.codebooks.shape == (4, 256, 8)
    pq.nprobe = 8
    D, I = pq.search(q, 30)
    assert (np.diff(D, axis=1) >= 0).all()
    decoded = pq.centroids[pq.assign] + np.concatenate(
        [pq.codebooks[s][pq.codes[:, s]] for s in range(4)], axis=1)
    for qi in range(len(q)):
        want = ((q[qi][None, :] - decoded[I[qi]]) ** 2).sum(1)
        assert np.allclose(D[qi], want, rtol=1e-5, atol=1e-6)
        # nothing outside the result beats the k-th distance (all lists probed)
        full = ((q[qi][None, :] - decoded) ** 2).sum(1)
        assert np.sort(full)[29] <= D[qi, 29] + 1e-5
    # the encoder picks the nearest codeword of every residual sub-vector
    r = x - pq.centroids[pq.assign]
    for s in range(4):
        d2 = ((r[:50, None, s * 8:(s + 1) * 8] - pq.codebooks[s][None]) ** 2).sum(-1)
        assert np.array_equal(d2.argmin(1), pq.codes[:50, s])


def test_oracle_hnsw_stand_in_is_exact_l2():
    """'HNSW' (faiss_retrieval.py:65-70, L2) is checked against the exact L2 answer it approximates: on
    unit-norm rows that is the Flat IP ranking with distance 2 - 2<q,x>; ascending; empty slots +FLT_MAX."""
    from oracle.flat import OracleFAISSIndex
    rng = np.random.default_rng(12)
    x = rng.standard_normal((700, 24)).astype(np.float32)
    q = rng.standard_normal((5, 24)).astype(np.float32)
    h, f = OracleFAISSIndex(24, 'HNSW'), OracleFAISSIndex(24, 'Flat')
    h.add(x), f.add(x)
    ih, dh = h.search(q, k=710)
    i_f, df = f.search(q, k=710)
    assert np.array_equal(ih[:, :700], i_f[:, :700])
    assert np.allclose(dh[:, :700], 2 - 2 * df[:, :700], atol=2e-6)
    assert (np.diff(dh[:, :700], axis=1) >= 0).all()
    assert (dh[:, 700:] == np.float32(3.4028234663852886e38)).all() and (ih[:, 700:] == 699).all()  # id_map[-1]
    brute = ((q[:, None, :] / np.linalg.norm(q, axis=1)[:, None, None]
              - (x / np.linalg.norm(x, axis=1, keepdims=True))[None]) ** 2).sum(-1)
    assert np.array_equal(np.argsort(brute, axis=1, kind="stable")[:, :50], ih[:, :50])


def test_flat_oracles_agree_with_an_independent_bruteforce():
    """faiss is not installable here (the retrieval oracle is unpinned against it), so at least check the
    restatement against an independent third-party exact search: scikit-learn's brute-force
    NearestNeighbors.  On L2-normalised rows cosine distance = 1 - <q, x> (Flat IP ranking) and squared
    euclidean distance is what the 'HNSW' stand-in's oracle reports."""
    from sklearn.neighbors import NearestNeighbors
    from oracle.flat import OracleFAISSIndex, normalize_L2
    rng = np.random.default_rng(21)
    x = rng.standard_normal((3000, 48)).astype(np.float32)
    q = rng.standard_normal((17, 48)).astype(np.float32)
    k = 120
    flat, l2 = OracleFAISSIndex(48, 'Flat'), OracleFAISSIndex(48, 'HNSW')
    flat.add(x), l2.add(x)
    ids_ip, d_ip = flat.search(q, k=k)
    ids_l2, d_l2 = l2.search(q, k=k)
    xn, qn = normalize_L2(x.copy()), normalize_L2(q.copy())
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(xn.astype(np.float64))
    dist, idx = nn.kneighbors(qn.astype(np.float64))
    assert (idx == ids_ip).mean() > 0.999          # identical up to fp32-vs-fp64 near-ties
    assert np.allclose(1.0 - dist, d_ip, atol=2e-6)
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean").fit(xn.astype(np.float64))
    dist, idx = nn.kneighbors(qn.astype(np.float64))
    assert (idx == ids_l2).mean() > 0.999
    assert np.allclose(dist, d_l2, atol=4e-6)


def test_oracle_stage1_matches_the_reference_tower_golden():
    """End-to-end Stage 1 on the CPU oracle (oracle towers -> oracle Flat wrapper) vs tests/golden/stage1_cfg1.npz,
    whose embeddings came from the reference's own two_tower_model.py (make_stage1_golden.py)."""
    from oracle.compare import compare_topk
    from oracle.flat import OracleFAISSIndex
    fx = np.load(GOLDEN / "stage1_cfg1.npz")
    cfg = CONFIGS["cfg1"]
    state = make_state(cfg, int(fx["state_seed"]))
    _, _, acat = make_inputs(cfg, int(fx["ad_seed"]), int(fx["n_ads"]))
    ucat, unum, _ = make_inputs(cfg, int(fx["user_seed"]), int(fx["n_users"]))
    u = otowers.tower_forward(state, "user_tower", ucat, unum)
    np.testing.assert_allclose(u, fx["user_out"], rtol=0, atol=5e-6)
    index = OracleFAISSIndex(cfg["output_dim"], 'Flat')
    index.add(otowers.tower_forward(state, "ad_tower", acat), [10 * i + 3 for i in range(int(fx["n_ads"]))])
    k = int(fx["k"])
    ids, dist = index.search(u, k=k)
    # the two tower implementations differ by ~1e-6 per component: order may differ inside 1e-5 score gaps only
    compare_topk(ids, dist, fx["ids"], fx["dist"], k, gap_tol=1e-5, score_rtol=0, score_atol=1e-5)


def test_oracle_wrapper_matches_the_reference_wrapper_run_over_a_faiss_stand_in():
    """tests/golden/wrapper_flat.npz holds what the reference's OWN `FAISSIndex` class (faiss_retrieval.py:14-256,
    executed unmodified by make_wrapper_golden.py over a numpy stand-in for the four faiss calls of the Flat path)
    returns: float64 input, custom + default ids, k > ntotal (labels -1 -> id_map[-1]), return_distances=False,
    batch_search chunking, get_stats, the .metadata keys.  The oracle wrapper must reproduce all of it; the GPU
    wrapper is compared with the oracle wrapper by tests/test_flat_gpu.py::test_wrapper_semantics."""
    from oracle.flat import OracleFAISSIndex
    fx = np.load(GOLDEN / "wrapper_flat.npz")
    o = OracleFAISSIndex(64, 'Flat')
    x1, x2, q = fx["x1"], fx["x2"], fx["q"]
    x1c, x2c, qc = x1.copy(), x2.copy(), q.copy()
    o.add(x1, ad_ids=fx["ids1"].tolist())
    o.add(x2)
    assert o.id_map == fx["id_map"].tolist() == fx["metadata_id_map"].tolist()
    ids, dist = o.search(q[:9], k=10)
    assert ids.dtype == fx["ids_k10"].dtype and np.array_equal(ids, fx["ids_k10"])
    np.testing.assert_allclose(dist, fx["dist_k10"], rtol=0, atol=2e-6)
    ids, dist = o.search(q[:3], k=520)
    assert np.array_equal(ids, fx["ids_big"])                      # the 20 missing slots carry id_map[-1]
    assert (ids[:, 500:] == o.id_map[-1]).all() and (dist[:, 500:] == np.float32(-3.4028234663852886e38)).all()
    np.testing.assert_allclose(dist[:, :500], fx["dist_big"][:, :500], rtol=0, atol=2e-6)
    assert np.array_equal(dist[:, 500:], fx["dist_big"][:, 500:])
    assert np.array_equal(o.search(q[:4], k=6, return_distances=False), fx["ids_only"])
    b_ids, b_dist = o.batch_search(q, k=5, batch_size=7)
    assert np.array_equal(b_ids, fx["batch_ids"])
    np.testing.assert_allclose(b_dist, fx["batch_dist"], rtol=0, atol=2e-6)
    stats = o.get_stats()
    assert list(stats.keys()) == fx["stats_keys"].tolist()
    assert [str(v) for v in stats.values()] == fx["stats_values"].tolist()
    assert fx["metadata_keys"].tolist() == ['dimension', 'index_type', 'nlist', 'nprobe', 'id_map']
    assert np.array_equal(x1, x1c) and np.array_equal(x2, x2c) and np.array_equal(q, qc)   # inputs never mutated


def test_stage1_fixture_is_what_the_reference_pipeline_returns():
    """stage1_cfg1.npz (reference towers + the ORACLE wrapper; the fixture the B200 path is compared with end to end)
    against stage1_ref_pipeline.npz: the same system run entirely through the reference's code - its towers, its
    FAISSIndex corpus build and its TwoStageRetriever.retrieve_and_rank, one user per call - over the numpy stand-in
    for faiss's IndexFlatIP / normalize_L2 (make_stage1_ref_pipeline.py)."""
    from oracle.compare import compare_topk
    fx, ref = np.load(GOLDEN / "stage1_cfg1.npz"), np.load(GOLDEN / "stage1_ref_pipeline.npz")
    k = int(fx["k"])
    assert int(ref["k"]) == k and ref["ids"].shape == (int(fx["n_users"]), k)
    # per-user (batch 1) towers vs the batched ones, sgemv vs sgemm scores: ~1e-7 apart, so order may differ
    # inside 1e-6 gaps only
    res = compare_topk(ref["ids"], ref["dist"], fx["ids"], fx["dist"], k, gap_tol=1e-6, score_rtol=0, score_atol=2e-6)
    assert res["exact_positions"] >= 0.99 * ref["ids"].size
