"""Parity comparators (SURVEY.md §8c "Parity criteria").

compare_topk: ids and order must be IDENTICAL to the oracle wherever adjacent oracle score
gaps exceed `gap_tol`; inside a run of near-ties (gap <= gap_tol) only the id SET must match.
The run that touches position k-1 may extend past k in the oracle (pass `extra` oracle
results) — there the GPU ids must be a subset of the extended run.
"""
from __future__ import annotations

from collections import Counter

import numpy as np


def compare_topk(ids, dist, ref_ids, ref_dist, k, *, gap_tol=1e-6, score_rtol=1e-3, score_atol=2e-6,
                 descending=True):
    """ref_ids/ref_dist hold k+extra columns (extra >= 0). Returns a dict of counters; raises
    AssertionError with a precise message on the first violation."""
    ids = np.asarray(ids)
    dist = np.asarray(dist, dtype=np.float32)
    ref_ids = np.asarray(ref_ids)
    ref_dist = np.asarray(ref_dist, dtype=np.float32)
    Q = ids.shape[0]
    assert ids.shape == (Q, k) and dist.shape == (Q, k), (ids.shape, dist.shape, k)
    sign = 1.0 if descending else -1.0
    exact_pos = tie_pos = 0
    max_abs = 0.0
    for q in range(Q):
        rd = ref_dist[q].astype(np.float64) * sign
        ri = ref_ids[q]
        n_valid = int((ri >= 0).sum()) if np.issubdtype(ri.dtype, np.integer) else len(ri)
        kk = min(k, n_valid)
        # scores
        err = np.abs(dist[q, :kk].astype(np.float64) - ref_dist[q, :kk].astype(np.float64))
        tol = score_atol + score_rtol * np.abs(ref_dist[q, :kk].astype(np.float64))
        if (err > tol).any():
            j = int(np.argmax(err - tol))
            raise AssertionError(f"query {q} pos {j}: score {dist[q, j]!r} vs oracle {ref_dist[q, j]!r}")
        if kk:
            max_abs = max(max_abs, float(err.max()))
        # runs of near-ties in the oracle list
        j = 0
        total = min(len(ri), n_valid)
        while j < kk:
            e = j
            while e + 1 < total and (rd[e] - rd[e + 1]) <= gap_tol:
                e += 1
            if e == j:
                if ids[q, j] != ri[j]:
                    raise AssertionError(
                        f"query {q} pos {j}: id {ids[q, j]} != oracle {ri[j]} "
                        f"(gaps {rd[j-1]-rd[j] if j else None}, {rd[j]-rd[j+1] if j+1 < total else None})")
                exact_pos += 1
            else:
                hi = min(e, kk - 1)
                got = Counter(ids[q, j:hi + 1].tolist())
                allowed = Counter(ri[j:e + 1].tolist())
                if e < kk:
                    ok = got == allowed
                else:  # the run extends past k: multiset inclusion
                    ok = all(allowed.get(key, 0) >= cnt for key, cnt in got.items())
                if not ok:
                    raise AssertionError(f"query {q} positions {j}..{hi}: ids {sorted(got)} not in near-tie run "
                                         f"{sorted(allowed)}")
                tie_pos += hi + 1 - j
            j = e + 1
        # empty slots
        if kk < k:
            assert (dist[q, kk:] == ref_dist[q, kk:k]).all(), f"query {q}: empty-slot distances differ"
            assert (ids[q, kk:] == ri[kk:k]).all(), f"query {q}: empty-slot ids differ"
    return {"exact_positions": exact_pos, "tie_positions": tie_pos, "max_abs_score_err": max_abs}


def recall_at_k(ids, truth_ids) -> float:
    """Mean fraction of each truth row found in the corresponding ids row."""
    ids = np.asarray(ids)
    truth_ids = np.asarray(truth_ids)
    hits = 0
    total = 0
    for a, t in zip(ids, truth_ids):
        t = t[t >= 0]
        hits += len(np.intersect1d(a, t))
        total += len(t)
    return hits / max(total, 1)
