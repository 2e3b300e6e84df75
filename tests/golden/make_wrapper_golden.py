"""Generates tests/golden/wrapper_flat.npz by running the UNMODIFIED reference wrapper
(`/root/reference/faiss_retrieval.py`, class FAISSIndex, index_type='Flat') in this container.

`faiss` itself cannot be imported here (no wheel, no network), so the four faiss entry points the wrapper touches
on the Flat path are provided by the tiny stand-in module below - `IndexFlatIP` (exact fp32 inner products, best
first, missing slots -1 / -FLT_MAX), `normalize_L2` (in place), `get_num_gpus` (0), `write_index` / `read_index`
(pickle).  Everything else that shapes the answer is the reference's own code, executed as is: the float32 copies,
the normalisation of corpus AND queries, the default-id continuation, the python `id_map[idx]` remap (with its
negative-index wrap for the -1 labels of k > ntotal), the (ad_ids, distances) return order, `batch_search`'s
chunking and vstack, `get_stats`, the `.metadata` side-car.  The fixture therefore pins the WRAPPER semantics to
the reference; the arithmetic of faiss's IndexFlatIP stays a restatement (two independent ones: the stand-in here,
written as a plain per-query stable argsort, and oracle/flat.py's partition + lexsort - the CPU test makes them meet).

Run in the build container only: `python tests/golden/make_wrapper_golden.py`.
"""
import pickle
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def _install_faiss_stand_in():
    m = types.ModuleType("faiss")
    m.METRIC_INNER_PRODUCT = 0
    m.METRIC_L2 = 1

    class IndexFlatIP:
        def __init__(self, d):
            self.d, self.is_trained = d, True
            self._x = np.zeros((0, d), dtype=np.float32)

        @property
        def ntotal(self):
            return len(self._x)

        def add(self, x):
            assert x.dtype == np.float32 and x.flags.c_contiguous and x.shape[1] == self.d
            self._x = np.vstack([self._x, x])

        def search(self, q, k):
            assert q.dtype == np.float32
            D = np.full((len(q), k), -3.4028234663852886e38, dtype=np.float32)
            I = np.full((len(q), k), -1, dtype=np.int64)
            for i, row in enumerate(q):
                s = self._x @ row
                order = np.argsort(-s, kind="stable")[:k]
                D[i, :len(order)] = s[order]
                I[i, :len(order)] = order
            return D, I

    def normalize_L2(x):
        assert x.dtype == np.float32 and x.ndim == 2
        for r in x:                      # in place, row by row; zero rows stay zero
            n2 = np.float32(np.dot(r, r))
            if n2 > 0:
                r *= np.float32(1.0) / np.sqrt(n2, dtype=np.float32)

    m.IndexFlatIP = IndexFlatIP
    m.normalize_L2 = normalize_L2
    m.get_num_gpus = lambda: 0
    m.write_index = lambda index, path: np.save(open(path, "wb"), index._x)      # any bytes will do: not compared
    m.read_index = lambda path: None
    sys.modules["faiss"] = m
    return m


def main():
    _install_faiss_stand_in()
    sys.path.insert(0, "/root/reference")
    import faiss_retrieval as ref      # the reference, untouched

    rng = np.random.default_rng(7301)
    d, n1, n2 = 64, 300, 200
    x1 = (rng.standard_normal((n1, d)) * rng.uniform(0.2, 5.0, (n1, 1))).astype(np.float64)   # float64 in: astype copy
    x2 = rng.standard_normal((n2, d)).astype(np.float32)
    x2[17] = 0.0                                                                            # a zero row stays zero
    ids1 = [5000 + 3 * i for i in range(n1)]
    q = rng.standard_normal((23, d)).astype(np.float32) * 3.0
    x1_before, x2_before, q_before = x1.copy(), x2.copy(), q.copy()

    idx = ref.FAISSIndex(dimension=d, index_type='Flat')
    idx.add(x1, ad_ids=ids1)
    idx.add(x2)                         # default ids continue from len(id_map)
    ids_k10, dist_k10 = idx.search(q[:9], k=10)
    ids_big, dist_big = idx.search(q[:3], k=n1 + n2 + 20)        # k > ntotal: -1 labels -> id_map[-1]
    ids_only = idx.search(q[:4], k=6, return_distances=False)
    b_ids, b_dist = idx.batch_search(q, k=5, batch_size=7)
    stats = idx.get_stats()
    assert np.array_equal(x1, x1_before) and np.array_equal(x2, x2_before) and np.array_equal(q, q_before)
    with tempfile.TemporaryDirectory() as tmp:
        idx.save(str(Path(tmp) / "sub" / "faiss_index.bin"))
        meta = pickle.load(open(Path(tmp) / "sub" / "faiss_index.bin.metadata", "rb"))
    np.savez_compressed(
        HERE / "wrapper_flat.npz", x1=x1, x2=x2, ids1=np.array(ids1), q=q,
        id_map=np.array(idx.id_map), ids_k10=ids_k10, dist_k10=dist_k10, ids_big=ids_big, dist_big=dist_big,
        ids_only=ids_only, batch_ids=b_ids, batch_dist=b_dist,
        stats_keys=np.array(list(stats.keys())), stats_values=np.array([str(v) for v in stats.values()]),
        metadata_keys=np.array(list(meta.keys())), metadata_id_map=np.array(meta["id_map"]))
    print("ids_k10", ids_k10.shape, ids_k10.dtype, "big", ids_big.shape, "tail id", ids_big[0, -1], "stats", stats)


if __name__ == "__main__":
    main()
