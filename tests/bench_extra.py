"""Secondary measurements for BASELINE configs 2 (batch sweep), 3 (IVF / IVF-PQ, 10M ads) and 4 (user-tower
encode, 26 x 10M-row tables, batch 65536).  One JSON line per measurement.  Not the driver's bench.py.
    python tests/bench_extra.py [sweep] [ivf] [ivfpq] [tower] [--small]
"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex, IndexFlatIP

FAISSIndex.verbose = False
PEAK_HBM, PEAK_TC = 6446.9, 1392.7
SMALL = "--small" in sys.argv
dev = torch.device("cuda", 0)


def timed(fn, warmup=3, steps=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def emit(**kw):
    print(json.dumps(kw), flush=True)


def sweep():
    N, d, k = 1_000_000, 256, 500
    g = torch.Generator(device=dev).manual_seed(1)
    idx = IndexFlatIP(d)
    idx.add(torch.randn((N, d), generator=g, device=dev), normalize=True)
    for Q in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
        q = torch.randn((Q, d), generator=g, device=dev)
        ms = timed(lambda: idx.search_device(q, k, normalize=True))
        idx.set_param("profile", 5)
        for _ in range(5):
            _, _, st, _ = idx.search_device(q, k, normalize=True)
        torch.cuda.synchronize()
        scan_ms = idx.get_param("scan_ms_avg")
        idx.set_param("profile", 0)
        hbm = (N * d * 2) / scan_ms / 1e6
        tf = 2.0 * Q * N * d / scan_ms / 1e9
        emit(cfg="flat_sweep", N=N, Q=Q, k=k, ms_per_step=ms, qps=Q / ms * 1e3, scan_kernel_ms=scan_ms,
             scan_GBps=hbm, scan_frac_hbm=hbm / PEAK_HBM, scan_TFLOPs=tf, scan_frac_tensor=tf / PEAK_TC,
             not_exact=int((st != 0).sum()))


def _mog(n, d, ncl, seed, chunk=1 << 20, centre_seed=3):
    centres = torch.randn((ncl, d), generator=torch.Generator(device=dev).manual_seed(centre_seed), device=dev)
    g = torch.Generator(device=dev).manual_seed(1000 + seed)
    out = torch.empty((n, d), device=dev)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        lab = torch.randint(0, ncl, (hi - lo,), generator=g, device=dev)
        out[lo:hi] = centres[lab] + 0.35 * torch.randn((hi - lo, d), generator=g, device=dev)
    return torch.nn.functional.normalize(out, dim=1)


def _recall(ids, truth):
    hits = 0
    for a, t in zip(ids, truth):
        hits += len(np.intersect1d(a, t))
    return hits / truth.size


def ivf(kind="IVF"):
    N, d, k = (1_000_000 if SMALL else 10_000_000), 256, 500
    nlist, nprobe = (1024 if SMALL else 4096), 32
    x = _mog(N, d, nlist, seed=3)
    qs = _mog(4096, d, nlist, seed=4)   # queries drawn from the same mixture as the data (different samples)
    t0 = time.time()
    idx = FAISSIndex(d, kind, nlist=nlist, nprobe=nprobe, pq_m=32)
    idx.train(x)
    torch.cuda.synchronize()
    t_train = time.time() - t0
    t0 = time.time()
    idx.add(x)
    torch.cuda.synchronize()
    t_add = time.time() - t0
    import os
    if os.environ.get("B2R_SKIP_TRUTH"):
        truth = None
    else:
        flat = IndexFlatIP(d)
        flat.add(x, normalize=True)
        _, truth = flat.search(qs[:512], k, normalize=True)
        del flat
        torch.cuda.empty_cache()
    sizes = idx.index.list_sizes()
    for Q in (1, 64, 4096):
        q = qs[:Q].contiguous()
        ms = timed(lambda: idx.index.search_device(q, k, normalize=True), steps=5)
        rec = None
        if truth is not None:
            ids, _ = idx.search(qs[:min(Q, 512)], k=k)
            rec = _recall(ids, truth[:min(Q, 512)])
        per_q_rows = nprobe * N / nlist
        emit(cfg=kind.lower(), N=N, nlist=nlist, nprobe=nprobe, pq_m=32 if kind == "IVFPQ" else None, Q=Q, k=k,
             ms_per_step=ms, qps=Q / ms * 1e3, recall_at_500_vs_flat=rec, train_s=t_train, add_s=t_add,
             list_size_min=int(sizes.min()), list_size_max=int(sizes.max()), rows_scanned_per_query=per_q_rows,
             not_exact=int((idx.index.last_status != 0).sum()) if idx.index.last_status is not None else None)


def tower():
    from movie_recommender_demo_b200.two_tower_model import UserTower
    rows = 1_000_000 if SMALL else 10_000_000
    torch.manual_seed(5)
    dims = {f"C{i + 1}": rows for i in range(26)}
    t = UserTower(dims, 13)
    for m in t.mlp:
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    t = t.to(dev).eval()
    t.check_indices = False
    t.embedding_layer.check_indices = False
    B = 65536
    g = torch.Generator(device=dev).manual_seed(6)
    cat = torch.randint(0, rows, (B, 26), generator=g, device=dev)
    num = torch.randn((B, 13), generator=g, device=dev)
    for a in sys.argv:                       # A/B of the fused kernel on CTA pairs: --pair=0,1,0,1
        if a.startswith("--pair="):
            for mode in a.split("=")[1].split(","):
                t.pair = int(mode)
                t._free()
                emit(cfg="user_tower_pair_ab", pair=int(mode), ms=timed(lambda: t(cat, num)))
            t.pair = None
            t._free()
    ms = timed(lambda: t(cat, num))
    ms_g = timed(lambda: t.embedding_layer(cat))
    flops = 2.0 * B * (429 * 512 + 512 * 256 + 256 * 256)
    gather_bytes = B * 26 * (64 + 8) + B * 26 * 64
    emit(cfg="user_tower", tables_rows=rows, table_GB=26 * rows * 64 / 1e9, B=B, ms=ms, users_per_s=B / ms * 1e3,
         tower_TFLOPs=flops / ms / 1e9, gather_only_ms=ms_g, gather_GBps=gather_bytes / ms_g / 1e6,
         gather_frac_hbm=gather_bytes / ms_g / 1e6 / PEAK_HBM)
    # zipf-distributed ids (real Criteo is heavy tailed)
    z = torch.from_numpy(np.random.default_rng(7).zipf(1.05, size=(B, 26)) % rows).to(dev)
    ms_z = timed(lambda: t(z, num))
    emit(cfg="user_tower_zipf", B=B, ms=ms_z, users_per_s=B / ms_z * 1e3)


def ranker():
    """Stage-2 ranker (inference.py:250-255 shape): 500 candidate rows of one user, and 64 users x 500 rows."""
    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    from weights import RANKER_CONFIGS, feature_dims, make_ranker_inputs, make_ranker_state
    from movie_recommender_demo_b200.transformer_ranker import TransformerRanker
    cfg = RANKER_CONFIGS["cfg1"]
    user, ad = feature_dims(cfg)
    m = TransformerRanker(user, ad, cfg["numerical_dim"], embedding_dim=cfg["embedding_dim"], d_model=cfg["d_model"],
                          num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], d_ff=cfg["d_ff"])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in make_ranker_state(cfg, 1).items()})
    m = m.to(dev).eval()
    per_row = 2.0 * (845 * 256 + 3 * (256 * 256 + 2 * 256 * 1024) + 3 * 256 * 256 + 3 * (256 * 256 + 256 * 64 + 64))
    for B in (500, 32000):
        ucat, acat, num = (torch.from_numpy(a).to(dev) for a in make_ranker_inputs(cfg, 2, B))
        with torch.no_grad():
            ms = timed(lambda: m(ucat, acat, num))
        emit(cfg="stage2_ranker", rows=B, ms=ms, rows_per_s=B / ms * 1e3, TFLOPs=per_row * B / ms / 1e9)


if __name__ == "__main__":
    what = [a for a in sys.argv[1:] if not a.startswith("--")] or ["sweep", "ivf", "ivfpq", "tower"]
    if "sweep" in what:
        sweep()
    if "tower" in what:
        tower()
    if "ivf" in what:
        ivf("IVF")
    if "ivfpq" in what:
        ivf("IVFPQ")
    if "ranker" in what:
        ranker()
