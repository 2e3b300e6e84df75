"""Property tests (hypothesis) of the oracle and the comparator: CPU only, seconds."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle.compare import compare_topk
from oracle.flat import NEG_FLT_MAX, OracleIndexFlatIP, normalize_L2, topk_desc


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 300), d=st.sampled_from([4, 16, 64]), q=st.integers(1, 5), k=st.integers(1, 40),
       seed=st.integers(0, 2**31 - 1), dup=st.booleans())
def test_topk_desc_is_the_stable_argsort_prefix(n, d, q, k, seed, dup):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if dup and n > 3:
        x[n // 2] = x[0]                                  # exact ties -> canonical order by label
    qq = rng.standard_normal((q, d)).astype(np.float32)
    S = qq @ x.T
    D, I = topk_desc(S, k)
    ref = np.argsort(-S, axis=1, kind="stable")[:, :k]
    take = min(k, n)
    assert np.array_equal(I[:, :take], ref[:, :take])
    assert (I[:, take:] == -1).all() and (D[:, take:] == NEG_FLT_MAX).all()
    assert (np.diff(D[:, :take], axis=1) <= 0).all()
    # the comparator accepts the oracle against itself and rejects a rotated row when gaps are real
    D2, I2 = topk_desc(S, k, extra=4)
    compare_topk(I, D, I2, D2, k)


@settings(max_examples=30, deadline=None)
@given(n=st.integers(2, 200), d=st.sampled_from([8, 32]), seed=st.integers(0, 2**31 - 1), scale=st.floats(1e-3, 1e3))
def test_normalize_l2_properties(n, d, seed, scale):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, d)) * scale).astype(np.float32)
    x[0] = 0
    y = normalize_L2(x.copy())
    norms = np.linalg.norm(y.astype(np.float64), axis=1)
    assert norms[0] == 0                                 # zero rows stay zero
    assert np.allclose(norms[1:], 1.0, atol=1e-6)
    assert np.allclose(normalize_L2(y.copy()), y, atol=2e-7)   # idempotent up to an ulp
    # scale invariance of the retrieval result: normalised search ignores the input scale
    idx = OracleIndexFlatIP(d)
    idx.add(y)
    q = rng.standard_normal((3, d)).astype(np.float32)
    _, I1 = idx.search(normalize_L2(q.copy()), min(5, n))
    _, I2 = idx.search(normalize_L2((q * np.float32(7.5)).copy()), min(5, n))
    assert np.array_equal(I1, I2)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(40, 400), parts=st.integers(2, 5), k=st.integers(1, 30), seed=st.integers(0, 2**31 - 1))
def test_sharded_merge_equals_unsharded(n, parts, k, seed):
    """shard -> local top-k -> pool -> top-k == unsharded top-k (the invariant the GPU merge relies on)."""
    from movie_recommender_demo_b200.sharded import shard_rows
    rng = np.random.default_rng(seed)
    x = normalize_L2(rng.standard_normal((n, 16)).astype(np.float32))
    q = normalize_L2(rng.standard_normal((4, 16)).astype(np.float32))
    S = q @ x.T
    Dref, Iref = topk_desc(S, k)
    pooled_D, pooled_I = [], []
    for r in range(parts):
        lo, hi = shard_rows(n, parts, r)
        D, I = topk_desc(S[:, lo:hi], k)
        pooled_D.append(D)
        pooled_I.append(np.where(I >= 0, I + lo, -1))
    PD, PI = np.concatenate(pooled_D, 1), np.concatenate(pooled_I, 1)
    Dm, pos = topk_desc(np.where(PI >= 0, PD, -np.inf).astype(np.float32), k)
    Im = np.take_along_axis(PI, pos, axis=1)
    take = min(k, n)
    assert np.array_equal(Im[:, :take], Iref[:, :take])
