"""Staged on-GPU diagnostics (not a pytest module): each stage runs in its own process so a
device trap in one stage cannot poison the next.  `python tests/gpu_diag.py [stage ...]`
writes gpurun_out/diag_<stage>.log and prints a one-line verdict per stage.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def _bf16_round(a):
    import torch
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


def stage_ingest():
    import numpy as np
    from movie_recommender_demo_b200.faiss_retrieval import IndexFlatIP
    from oracle.flat import normalize_L2
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 256)).astype(np.float32) * 3
    x[5] = 0
    idx = IndexFlatIP(256)
    idx.add(x, normalize=True)
    got = idx.reconstruct_n(0, 1000).cpu().numpy()
    ref = normalize_L2(x.copy())
    err = np.abs(got - ref).max()
    print("ntotal", idx.ntotal, "max err vs oracle normalize", err, "zero row stays zero", (got[5] == 0).all())
    assert err < 2e-7 and (got[5] == 0).all()
    idx.add(x[:10], normalize=False)
    got2 = idx.reconstruct_n(1000, 10).cpu().numpy()
    assert (got2 == x[:10]).all(), "un-normalised add must be a bit-exact copy"
    print("OK")


def _check_scores(N, Q, d, seed=0, structured=False):
    import numpy as np
    import torch
    from movie_recommender_demo_b200.faiss_retrieval import IndexFlatIP
    rng = np.random.default_rng(seed)
    if structured:
        x = np.zeros((N, d), np.float32)
        x[np.arange(N), np.arange(N) % d] = 1.0 + (np.arange(N) // d)
        q = np.zeros((Q, d), np.float32)
        for i in range(Q):
            q[i, (3 * i) % d] = 1.0
            q[i, (3 * i + 1) % d] = 0.5
    else:
        x = rng.standard_normal((N, d)).astype(np.float32)
        q = rng.standard_normal((Q, d)).astype(np.float32)
    idx = IndexFlatIP(d)
    idx.add(x, normalize=False)
    tc = idx.debug_scores(q, "tc").cpu().numpy()
    torch.cuda.synchronize()
    simt = idx.debug_scores(q, "simt").cpu().numpy()
    ref = _bf16_round(q).astype(np.float64) @ _bf16_round(x).astype(np.float64).T
    e_tc = np.abs(tc - ref).max()
    e_simt = np.abs(simt - ref).max()
    scale = np.abs(ref).max()
    print(f"N={N} Q={Q} d={d} structured={structured}: |tc-ref|={e_tc:.3e} |simt-ref|={e_simt:.3e} scale={scale:.3e}")
    if e_tc > 1e-3 * max(scale, 1):
        OUT.mkdir(exist_ok=True)
        np.save(OUT / f"bad_tc_{N}_{Q}_{d}_{int(structured)}.npy", tc[:256, :1024])
        np.save(OUT / f"bad_ref_{N}_{Q}_{d}_{int(structured)}.npy", ref[:256, :1024].astype(np.float32))
        bad = np.argwhere(np.abs(tc - ref) > 1e-3 * max(scale, 1))
        print("  first mismatches (q,row):", bad[:10].tolist(), "count", len(bad), "of", tc.size)
        return False
    return True


def stage_scores_small():
    ok = _check_scores(256, 16, 256, structured=True)
    ok &= _check_scores(256, 16, 256)
    ok &= _check_scores(128, 1, 64)
    assert ok
    print("OK")


def stage_scores_sizes():
    ok = True
    for (N, Q, d) in [(1000, 1, 256), (5000, 130, 256), (70001, 300, 128), (100000, 64, 64), (33, 5, 192),
                      (300000, 257, 256)]:
        ok &= _check_scores(N, Q, d, seed=N)
    assert ok
    print("OK")


def _search_case(N, Q, k, d=256, force_path=0, seed=0, tag=""):
    import numpy as np
    import torch
    from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
    from oracle.flat import OracleFAISSIndex
    from oracle.compare import compare_topk
    FAISSIndex.verbose = False
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, d)).astype(np.float32)
    q = rng.standard_normal((Q, d)).astype(np.float32)
    g = FAISSIndex(d, 'Flat')
    if force_path:
        g.index.set_param("force_path", force_path)
    g.add(x)
    o = OracleFAISSIndex(d, 'Flat')
    o.add(x)
    t0 = time.time()
    ids, dist = g.search(q, k=k)
    torch.cuda.synchronize()
    t1 = time.time()
    rid, rd = o.search(q, k=k, extra=32)
    res = compare_topk(ids, dist, rid, rd, k)
    print(f"{tag} N={N} Q={Q} k={k} d={d} force={force_path}: {res} retries={g.index.last_retries} "
          f"status_nonzero={(g.index.last_status != 0).sum()} gpu_s={t1 - t0:.4f}")


def stage_search_dense():
    _search_case(20000, 37, 50, tag="dense")
    _search_case(3000, 5, 500, tag="dense")
    _search_case(100, 3, 500, tag="k>N")
    _search_case(100000, 512, 500, tag="cfg1-shape")
    print("OK")


def stage_search_filter():
    _search_case(300000, 8, 100, force_path=2, tag="filter-forced")
    _search_case(1000000, 64, 500, tag="filter")
    _search_case(1000000, 1, 500, seed=3, tag="filter")
    _search_case(600000, 300, 500, seed=4, tag="filter-MQ2")
    print("OK")


def stage_timing():
    import numpy as np
    import torch
    from movie_recommender_demo_b200.faiss_retrieval import IndexFlatIP
    d, N = 256, 1_000_000
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((N, d), generator=g, device="cuda")
    idx = IndexFlatIP(d)
    idx.add(x, normalize=True)
    del x
    for Q in (1, 8, 64, 256, 1024, 4096):
        q = torch.randn((Q, d), generator=g, device="cuda")
        for _ in range(3):
            idx.search_device(q, 500, normalize=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            D, I, st, tr = idx.search_device(q, 500, normalize=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"Q={Q}: {ms:.3f} ms/search  {Q / ms * 1e3:.0f} q/s  status_nonzero={(st != 0).sum().item()} "
              f"scan_bytes_GBps={(N * d * 2) / ms / 1e6:.0f} tflops={2.0 * Q * N * d / ms / 1e9:.1f}")
    print("OK")


STAGES = {
    "ingest": stage_ingest,
    "scores_small": stage_scores_small,
    "scores_sizes": stage_scores_sizes,
    "search_dense": stage_search_dense,
    "search_filter": stage_search_filter,
    "timing": stage_timing,
}


def main():
    names = sys.argv[1:] or list(STAGES)
    if len(names) == 1 and names[0].startswith("--run="):
        STAGES[names[0][6:]]()
        return 0
    OUT.mkdir(exist_ok=True)
    rc_all = 0
    for name in names:
        log = OUT / f"diag_{name}.log"
        t0 = time.time()
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, __file__, f"--run={name}"], stdout=f, stderr=subprocess.STDOUT,
                                   timeout=420, cwd=str(ROOT))
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = -999
        tail = log.read_text().strip().splitlines()[-12:]
        print(f"[diag] {name}: rc={rc} ({time.time() - t0:.1f}s)")
        for line in tail:
            print("    " + line)
        rc_all |= (rc != 0)
    return rc_all


if __name__ == "__main__":
    sys.exit(main())
