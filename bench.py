#!/usr/bin/env python
"""bench.py — Stage-1 retrieval throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--batch Q] [--impl reference]

metric  : retrieval queries/sec @ top-500 (Flat inner product, d=256)
N = 1   : BASELINE configs[1] — 1M x 256 ad corpus on one B200, query batch Q (default 4096)
N > 1   : BASELINE configs[4] — 100M x 256 corpus row-sharded over N ranks (strong scaling),
          every rank searches its shard for the same Q queries; the per-rank top-500 lists are exchanged
          (packed all-to-all by query slice + all-gather, or one packed all-gather for small batches) and
          merged by b2r_topk_merge_packed (movie_recommender_demo_b200/sharded.py).
A "step" = one pass of the hot path over one batch of Q queries:
  value : queries already resident in HBM, results left in HBM (CUDA-event timed, max over ranks)
  e2e   : FAISSIndex.search(numpy queries in pinned host memory) -> numpy ids + distances,
          host<->device copies and the id remap inside the timed region
  roofline    : the dominant kernel (filter scan, tcgen05) timed with CUDA events on its stream
  cpu_baseline: the reference-style CPU path (oracle port: fp32 sgemm + top-k + python id remap)
                on this box's host cores, on a bounded sample of the same workload
  parity_spot_check : N = 1: 64 queries drawn from every pipeline chunk vs the CPU oracle over the full corpus;
                N > 1 (and any corpus above 4M rows): every corpus chunk is REGENERATED from its seed
                (seed = 100 + chunk), scored in fp32 by the CPU oracle for 8 queries, merged, and compared
                (ids + order outside 1e-6 near-ties, scores) with what the sharded search returned
  extra : (N = 1 only) the other BASELINE configs measured in the same run — Flat at batch 1 and 64, a
          CLUSTERED 1M corpus (retry count reported), IVF-Flat / IVF-PQ over 10M ads, the user-tower encode at
          batch 65536 over 26 x 10M-row tables, and the 100M-row corpus on ONE GPU (the strong-scaling anchor
          of the N > 1 lines); each with its own value and roofline.  --no-extra skips them.
`--impl reference` times only the CPU path (the reference's faiss-cpu wheel is not installable here: no
network, no wheel — see DESIGN.md), same config / metric / unit: at N = 1 the FULL query batch against the full
corpus (same_config true); at N > 1 a 1M-row slice scaled to the corpus size, labelled "extrapolated".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D = 256
K_TOP = 500
CORPUS_1GPU = 1_000_000
CORPUS_MULTI = 100_000_000
CHUNK = 1 << 20
METRIC = "retrieval queries/sec @ top-500 (Flat IP, d=256)"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run.  The period is 100 ms:
    at 20 ms the NVML polling itself slowed the launch- and copy-heavy e2e leg by ~30 % (measured)."""
    PERIOD_MS = 100
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", str(self.PERIOD_MS),
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def summary(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for r in self.rows if any(a <= r[0] <= b for a, b in windows)]
        if len(inside) < 3 and windows:  # short timed regions: everything from the first start to the last end
            pad = 1.5 * self.PERIOD_MS / 1e3
            inside = [r for r in self.rows if windows[0][0] - pad <= r[0] <= windows[-1][1] + pad]
        for ts, f in inside:
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self, windows):
        if self.proc is None:
            return self.summary(windows)
        time.sleep(0.15)
        self.proc.terminate()
        return self.summary(windows)


# ----------------------------------------------------------------------------- reference arm
def cpu_search_fn(x_host, id_map):
    """The reference-style CPU search: normalise (copy), fp32 sgemm, top-k best-first, python id remap
    (faiss_retrieval.py:146-160).  oracle port on torch-CPU/MKL; faiss itself is not available."""
    import numpy as np
    import torch
    from oracle.flat import normalize_L2
    xt = torch.from_numpy(x_host)

    def search(q, k):
        qn = q.astype('float32')
        normalize_L2(qn)
        S = torch.from_numpy(qn) @ xt.T
        Dv, Iv = torch.topk(S, k, dim=1, largest=True, sorted=True)
        indices = Iv.numpy()
        ad_ids = np.array([[id_map[idx] for idx in row] for row in indices])
        return ad_ids, Dv.numpy()
    return search


def cpu_search_batched(search, q, k, piece=512):
    """The full query batch through the CPU search, `piece` queries per call (the [piece, N] fp32 score matrix
    is 2 GB at N = 1M; the reference's wrapper has `batch_search` for exactly this, faiss_retrieval.py:168-194)."""
    import numpy as np
    ids, ds = [], []
    for lo in range(0, len(q), piece):
        a, b = search(q[lo:lo + piece], k)
        ids.append(a)
        ds.append(b)
    return np.vstack(ids), np.vstack(ds)


def make_host_corpus(n, seed=1):
    import numpy as np
    from oracle.flat import normalize_L2
    rng = np.random.default_rng(seed)
    x = np.empty((n, D), dtype=np.float32)
    for lo in range(0, n, CHUNK):
        hi = min(n, lo + CHUNK)
        x[lo:hi] = rng.standard_normal((hi - lo, D), dtype=np.float32)
    normalize_L2(x)
    return x


def run_reference(args):
    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    torch.set_num_threads(os.cpu_count() or 1)   # torchrun pins OMP_NUM_THREADS=1; use every host core
    world = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
    total_rows = args.corpus_rows or (CORPUS_1GPU if world == 1 else CORPUS_MULTI)
    # N = 1: the SAME configuration as the B200 arm - the full query batch against the full 1M-row corpus.
    # N > 1: the 100M x 256 fp32 corpus (102 GB) is not materialised on the host; a 1M-row slice is timed and
    # the throughput scaled by rows (a flat scan is linear in rows): an extrapolation, labelled as such.
    n = min(total_rows, CORPUS_1GPU)
    scale = n / total_rows
    x = make_host_corpus(n)
    search = cpu_search_fn(x, list(range(n)))
    rng = np.random.default_rng(2)
    q = rng.standard_normal((args.batch, D), dtype=np.float32)
    t0 = time.perf_counter()
    cpu_search_batched(search, q, K_TOP)                 # warm-up = one full step (also sizes the run)
    t_step = time.perf_counter() - t0
    budget = float(os.environ.get("B2R_REF_BUDGET_S", "150"))
    steps = max(1, min(args.steps, int(budget / max(t_step, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_search_batched(search, q, K_TOP)
    dt = time.perf_counter() - t0
    val = steps * args.batch / dt * scale
    cores = torch.get_num_threads()
    sample = (f"the full {args.batch}-query batch per step vs " + (f"the full {total_rows}x{D} corpus" if scale == 1 else
              f"a {n}-row slice of the {total_rows}x{D} corpus; queries/s scaled by {scale:g} to the full corpus")
              + f", top-{K_TOP}; {steps} timed steps of the {args.steps} requested (bounded to ~{budget:.0f} s of CPU work)")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "steps_requested": args.steps, "warmup": 1, "ms_per_step": dt / steps * 1e3 / scale,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "same_config": scale == 1, "extrapolated": scale != 1,
        "config": {"workload": (f"Flat IP top-{K_TOP} over {total_rows}x{D} ad corpus, query batch {args.batch}"
                                + (f", row-sharded over {world} B200 + packed all-to-all / all-gather merge" if world > 1
                                   else ", single B200") + " [CPU reference arm]"),
                   "corpus_rows": total_rows, "dim": D, "k": K_TOP, "batch": args.batch},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle port (torch-CPU sgemm + topk + python id remap); faiss-cpu is not installable here"},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------- helpers of the B200 arm
def _timed(torch, fn, warmup, steps, spin_up_s=0.03):
    """CUDA-event time per call.  Warm-up = `warmup` calls AND at least `spin_up_s` of them: the sub-results are
    measured right after host-side set-up during which the GPU idles and drops its clocks, and a handful of
    100-microsecond calls does not bring them back."""
    t0 = time.perf_counter()
    n = 0
    while n < warmup or time.perf_counter() - t0 < spin_up_s:
        fn()
        n += 1
        if n >= warmup:
            torch.cuda.synchronize()     # the clock below must see device time, not enqueue time
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _pick_tensor_peak(peaks, clocks, region_s):
    """Burst vs sustained cuBLAS figure: the sustained one only when the timed region is seconds long AND the
    sampled SM clock sat visibly below the maximum (a power-capped steady state); else the burst figure."""
    sm, smax = (clocks or {}).get("sm_mhz"), (clocks or {}).get("sm_max_mhz")
    capped = bool(sm and smax and sm < 0.9 * smax)
    if region_s >= 1.0 and capped:
        return peaks["bf16_tflops_sustained"], f"{peaks['source']} cuBLAS bf16, sustained (timed region {region_s:.1f} s at {sm:.0f}/{smax:.0f} MHz)"
    return peaks["bf16_tflops"], (f"{peaks['source']} cuBLAS bf16, burst (timed region {region_s * 1e3:.0f} ms"
                                  + (f" at {sm:.0f}/{smax:.0f} MHz)" if sm and smax else ")"))


def _mog(torch, dev, n, ncl, seed, centre_seed=3):
    """unit-norm mixture of `ncl` Gaussians generated on the device (SURVEY §8d cfg 3)."""
    centres = torch.randn((ncl, D), generator=torch.Generator(device=dev).manual_seed(centre_seed), device=dev)
    g = torch.Generator(device=dev).manual_seed(1000 + seed)
    out = torch.empty((n, D), device=dev)
    for lo in range(0, n, CHUNK):
        hi = min(n, lo + CHUNK)
        lab = torch.randint(0, ncl, (hi - lo,), generator=g, device=dev)
        out[lo:hi] = centres[lab] + 0.35 * torch.randn((hi - lo, D), generator=g, device=dev)
    return torch.nn.functional.normalize(out, dim=1)


def _chunk_rows(torch, dev, gen, c):
    gen.manual_seed(100 + c)
    return torch.randn((CHUNK, D), generator=gen, device=dev)


def parity_by_regeneration(torch, dev, total_rows, q_np, ids, dist, nq=8):
    """SURVEY §8(d) cfg 5: the corpus is never on the host, so regenerate EVERY chunk from its seed, let the CPU
    oracle score it in fp32 for `nq` queries (exact top-(k+32) per chunk, canonical order), merge the per-chunk
    lists and compare with the (sharded) search's ids / order / scores."""
    import numpy as np
    from oracle.compare import compare_topk
    from oracle.flat import normalize_L2, topk_desc
    qn = q_np[:nq].astype(np.float32).copy()
    normalize_L2(qn)
    gen = torch.Generator(device=dev)
    keep = K_TOP + 32
    best_d = np.full((nq, 0), 0, dtype=np.float32)
    best_i = np.full((nq, 0), 0, dtype=np.int64)
    for c in range((total_rows + CHUNK - 1) // CHUNK):
        rows = _chunk_rows(torch, dev, gen, c)[: min(CHUNK, total_rows - c * CHUNK)].cpu().numpy()
        normalize_L2(rows)                              # faiss_retrieval.py:115 (the wrapper normalises on add)
        dc, ic = topk_desc(qn @ rows.T, K_TOP, extra=32)
        best_d = np.concatenate([best_d, dc], axis=1)
        best_i = np.concatenate([best_i, np.where(ic >= 0, ic + c * CHUNK, -1)], axis=1)
        if best_d.shape[1] > 8 * keep:                  # prune the pool now and then
            order = np.lexsort((best_i, -best_d), axis=1)[:, :keep]
            best_d, best_i = np.take_along_axis(best_d, order, 1), np.take_along_axis(best_i, order, 1)
    order = np.lexsort((best_i, -best_d), axis=1)[:, :keep]
    rd, rid = np.take_along_axis(best_d, order, 1), np.take_along_axis(best_i, order, 1)
    res = compare_topk(ids[:nq], dist[:nq], rid, rd, K_TOP, gap_tol=1e-6)
    return (f"pass ({nq} queries vs the CPU oracle over all {total_rows} regenerated rows: ids/order identical outside "
            f"1e-6 gaps, {res['exact_positions']} exact + {res['tie_positions']} near-tie positions, "
            f"max |score err| {res['max_abs_score_err']:.1e})")


# ------------------------------------------------------------------------------- extra configs (N = 1)
def run_extras(args, torch, dev, peaks, flat_index, q_dev, emit_err, skip=(), progress=None):
    """BASELINE configs other than the headline, one sub-result each (value + roofline).  `skip`: names not to
    run; `progress(name, extra)` is called before (name) and after (None) every sub-result so that the caller can
    persist what is finished."""
    import numpy as np
    from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex, IndexFlatIP
    extra = {}
    hbm = peaks["hbm_gbs"]

    def guarded(name, fn):
        if name in skip:
            return
        if progress:
            progress(name, extra)
        try:
            extra[name] = fn()
        except Exception as exc:  # report, never hide; the headline line must still be printed
            extra[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            emit_err(f"extra {name} failed: {exc!r}")
        torch.cuda.empty_cache()
        torch.cuda.synchronize()      # a sticky device error surfaces HERE, attributed to this sub-result
        if progress:
            progress(None, extra)

    # ---- configs[1] at the memory-bound batches
    def flat_small(Q):
        def run():
            idx = flat_index.index
            qs = q_dev[:Q].contiguous()
            ms = _timed(torch, lambda: idx.search_device_static(qs, K_TOP, normalize=True), 5, 50)
            idx.set_param("profile", 10)
            for _ in range(10):
                _, _, st, _ = idx.search_device(qs, K_TOP, normalize=True)
            torch.cuda.synchronize()
            scan_ms = idx.get_param("scan_ms_avg")
            idx.set_param("profile", 0)
            n = idx.ntotal
            kernel_bytes = n * D * 2.0 + Q * D * 2.0
            step_bytes = n * D * 2.0 + Q * (K_TOP * D * 4 + D * 4 + K_TOP * 12)
            q_np_small = q_dev[:Q].cpu().numpy()
            for _ in range(10):          # warm-up: pinned result blocks exist, clocks are up
                flat_index.search(q_np_small, k=K_TOP)
            t0 = time.perf_counter()
            for _ in range(50):
                flat_index.search(q_np_small, k=K_TOP)
            e2e_ms = (time.perf_counter() - t0) / 50 * 1e3
            return {"workload": f"Flat IP top-{K_TOP} over {n}x{D}, query batch {Q}, CUDA-graph replay",
                    "value": Q / ms * 1e3, "unit": "queries/s", "ms_per_step": ms,
                    "e2e": {"value": Q / e2e_ms * 1e3, "unit": "queries/s", "ms_per_step": e2e_ms},
                    "roofline": {"bound": "hbm", "achieved": kernel_bytes / scan_ms / 1e6, "peak": hbm, "unit": "GB/s",
                                 "frac": kernel_bytes / scan_ms / 1e6 / hbm, "kernel": "scan_tc_kernel<1,FILTER>",
                                 "kernel_ms": scan_ms,
                                 "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / ms / 1e6,
                                                "frac": step_bytes / ms / 1e6 / hbm}},
                    "queries_not_provably_exact": int((st != 0).sum())}
        return run
    guarded("flat_q1", flat_small(1))
    guarded("flat_q64", flat_small(64))

    # ---- a CLUSTERED corpus (tower outputs are clustered post-ReLU vectors; isotropic noise is the sampled
    #      threshold planner's best case): retry count and flagged queries reported
    def clustered():
        x = _mog(torch, dev, CORPUS_1GPU, 4096, seed=3)
        qs = _mog(torch, dev, args.batch, 4096, seed=4)
        idx = FAISSIndex(D, 'Flat', device=dev.index)
        idx.add(x)
        del x
        ms = _timed(torch, lambda: idx.index.search_device(qs, K_TOP, normalize=True), 3, 10)
        _, _, st, _ = idx.index.search_device(qs, K_TOP, normalize=True)
        flagged_dev = int((st != 0).sum())
        q_np = qs.cpu().numpy()
        idx.search(q_np, k=K_TOP)
        t0 = time.perf_counter()
        for _ in range(5):
            idx.search(q_np, k=K_TOP)
        e2e_ms = (time.perf_counter() - t0) / 5 * 1e3
        return {"workload": f"Flat IP top-{K_TOP} over a CLUSTERED {CORPUS_1GPU}x{D} corpus (4096-component mixture), "
                            f"query batch {args.batch} from the same mixture",
                "value": args.batch / ms * 1e3, "unit": "queries/s", "ms_per_step": ms,
                "e2e": {"value": args.batch / e2e_ms * 1e3, "unit": "queries/s", "ms_per_step": e2e_ms},
                "flagged_before_retry": flagged_dev, "retries_in_e2e_call": int(idx.index.last_retries),
                "queries_not_provably_exact_after_retry": int((idx.index.last_status != 0).sum())}
    guarded("flat_clustered", clustered)

    # ---- configs[2]: IVF-Flat / IVF-PQ over 10M ads
    def ivf(kind):
        def run():
            N, nlist, nprobe = 10_000_000, 4096, 32
            x = _mog(torch, dev, N, nlist, seed=3)
            qs = _mog(torch, dev, 4096, nlist, seed=4)
            idx = FAISSIndex(D, kind, nlist=nlist, nprobe=nprobe, pq_m=32, device=dev.index)
            t0 = time.time()
            idx.train(x)
            idx.add(x)
            torch.cuda.synchronize()
            build_s = time.time() - t0
            flat = IndexFlatIP(D, device=dev.index)
            flat.add(x, normalize=True)
            del x
            _, truth = flat.search(qs[:256], K_TOP, normalize=True)
            del flat
            torch.cuda.empty_cache()
            ids, _ = idx.search(qs[:256], k=K_TOP)
            rec = float(np.mean([len(np.intersect1d(a, t)) for a, t in zip(ids, truth)]) / K_TOP)
            per_q_bytes = nprobe * (N / nlist) * (32 if kind == "IVFPQ" else D * 2)
            out = {"workload": f"{kind} nlist {nlist} nprobe {nprobe}" + (" m=32 8-bit" if kind == "IVFPQ" else "")
                               + f" top-{K_TOP} over {N}x{D} (4096-component mixture)", "unit": "queries/s",
                   "recall_at_500_vs_flat": rec, "build_s": round(build_s, 2), "batches": {}}
            for Q in (1, 64, 4096):
                qq = qs[:Q].contiguous()
                ms = _timed(torch, lambda: idx.index.search_device(qq, K_TOP, normalize=True), 3, 10)
                _, _, st, _ = idx.index.search_device(qq, K_TOP, normalize=True)
                ach = Q * per_q_bytes / ms / 1e6
                out["batches"][str(Q)] = {"value": Q / ms * 1e3, "ms_per_step": ms,
                                          # fused list scan (IVF-Flat, Q >= 512): queries its sampled threshold left
                                          # flagged; FAISSIndex.search re-runs those through the dump path
                                          "flagged_before_fallback": int((st != 0).sum()),
                                          "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                                                       "frac": ach / hbm,
                                                       "note": "algorithmic bytes = Q x nprobe x (N/nlist) x bytes/row; "
                                                               "above 1 when queries share list reads (large batches)"}}
            out["value"] = out["batches"]["4096"]["value"]
            out["ms_per_step"] = out["batches"]["4096"]["ms_per_step"]
            return out
        return run
    guarded("ivf_10M", ivf("IVF"))
    guarded("ivfpq_10M", ivf("IVFPQ"))

    # ---- configs[3]: user-tower encode, 26 x 10M-row tables, 13 dense, batch 65536
    def tower():
        from movie_recommender_demo_b200.two_tower_model import UserTower
        rows, B = 10_000_000, 65536
        torch.manual_seed(5)
        t = UserTower({f"C{i + 1}": rows for i in range(26)}, 13)
        for m in t.mlp:
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        t = t.to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(6)
        cat = torch.randint(0, rows, (B, 26), generator=g, device=dev)
        num = torch.randn((B, 13), generator=g, device=dev)
        ms = _timed(torch, lambda: t(cat, num), 5, 20)
        t.check()
        ms_g = _timed(torch, lambda: t.embedding_layer(cat), 3, 10)
        one_c, one_n = cat[:1].contiguous(), num[:1].contiguous()
        ms1 = _timed(torch, lambda: t(one_c, one_n), 10, 200)
        flops = 2.0 * B * (429 * 512 + 512 * 256 + 256 * 256)
        nbytes = B * (26 * 64 + 26 * 8 + 13 * 4 + 256 * 4)
        tf = flops / ms / 1e9
        return {"workload": f"UserTower encode, 26 x {rows}-row fp32 tables (16.6 GB), 13 dense, 429->512->256->256, batch {B}",
                "value": B / ms * 1e3, "unit": "users/s", "ms_per_step": ms, "fused_kernel": True,
                "operand_dtype": t.native_operand_dtype, "batch1_latency_ms": ms1, "gather_only_ms": ms_g,
                "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": tf / peaks["bf16_tflops"],
                             "hbm": {"algorithmic_bytes": nbytes, "achieved": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / hbm}}}
    guarded("tower_65536", tower)

    # ---- the Stage-2 ranker on the 500 candidates of one user / of 64 users
    def ranker():
        sys.path.insert(0, str(ROOT / "tests" / "golden"))
        from weights import RANKER_CONFIGS, feature_dims, make_ranker_inputs, make_ranker_state
        from movie_recommender_demo_b200.transformer_ranker import TransformerRanker
        cfg = RANKER_CONFIGS["cfg1"]
        user, ad = feature_dims(cfg)
        m = TransformerRanker(user, ad, cfg["numerical_dim"], embedding_dim=cfg["embedding_dim"], d_model=cfg["d_model"],
                              num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], d_ff=cfg["d_ff"])
        m.load_state_dict({k: torch.from_numpy(v) for k, v in make_ranker_state(cfg, 1).items()})
        m = m.to(dev).eval()
        per_row = 2.0 * (845 * 256 + 3 * (256 * 256 + 2 * 256 * 1024) + 3 * 256 * 256 + 3 * (256 * 256 + 256 * 64 + 64))
        out = {"workload": "TransformerRanker (d_model 256, 3 layers, d_ff 1024, 3 cross layers, 3 heads) on stage-1 candidates",
               "unit": "rows/s", "batches": {}}
        for B in (K_TOP, 64 * K_TOP):
            ucat, acat, num = (torch.from_numpy(a).to(dev) for a in make_ranker_inputs(cfg, 2, B))
            with torch.no_grad():
                ms = _timed(torch, lambda: m(ucat, acat, num), 5, 20)
            out["batches"][str(B)] = {"value": B / ms * 1e3, "ms_per_step": ms, "TFLOPs": per_row * B / ms / 1e9}
        out["value"] = out["batches"][str(64 * K_TOP)]["value"]
        out["ms_per_step"] = out["batches"][str(64 * K_TOP)]["ms_per_step"]
        out["operand_dtype"] = m.native_operand_dtype
        return out
    guarded("stage2_ranker", ranker)

    # ---- the reference's production request (inference.py:199-288): ONE user dict on the host -> preprocess ->
    #      user tower -> top-500 over the 1M-row corpus -> Stage-2 ranker over the 500 candidates -> top-10 on the
    #      host.  Wall-clock per request through AdRecommenderInference.recommend_ads (encode + search end to end).
    def request():
        sys.path.insert(0, str(ROOT / "tests" / "golden"))
        from weights import CONFIGS, RANKER_CONFIGS, feature_dims, make_ranker_state, make_state
        from movie_recommender_demo_b200.inference import AdRecommenderInference
        from movie_recommender_demo_b200.transformer_ranker import TransformerRanker
        from movie_recommender_demo_b200.two_tower_model import TwoTowerModel
        cfg, rcfg = CONFIGS["cfg1"], RANKER_CONFIGS["cfg1"]
        user, ad = feature_dims(cfg)
        tt = TwoTowerModel(user, ad, cfg["numerical_dim"], cfg["embedding_dim"], cfg["hidden_dims"], cfg["output_dim"])
        tt.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in make_state(cfg, 1).items()})
        ruser, rad = feature_dims(rcfg)
        rk = TransformerRanker(ruser, rad, rcfg["numerical_dim"], embedding_dim=rcfg["embedding_dim"],
                               d_model=rcfg["d_model"], num_heads=rcfg["num_heads"], num_layers=rcfg["num_layers"],
                               d_ff=rcfg["d_ff"])
        rk.load_state_dict({k: torch.from_numpy(v) for k, v in make_ranker_state(rcfg, 1).items()})

        class _Enc:                      # duck-typed fitted LabelEncoder / StandardScaler of the reference's
            def __init__(self, n):       # CriteoDataPreprocessor (CPU ETL, out of scope)
                self.classes_ = np.array([f"v{j}" for j in range(n - 1)] + ["missing"], dtype=object)

        class _Pre:
            feature_dims = {**user, **ad}
            label_encoders = {c: _Enc(n) for c, n in {**user, **ad}.items()}
            numerical_cols = [f"I{i + 1}" for i in range(cfg["numerical_dim"])]

            class scaler:
                @staticmethod
                def transform(x):
                    return np.asarray(x, dtype=np.float32)

        users = [{"categorical": {f"C{i + 1}": f"v{(7 * u + i) % 40}" for i in range(6)},
                  "numerical": {f"I{i + 1}": float(u % 17 + i) for i in range(cfg["numerical_dim"])}} for u in range(64)]
        out = {"workload": f"AdRecommenderInference.recommend_ads, one request: host dict -> user tower -> Flat top-{K_TOP} "
                           f"over {flat_index.index.ntotal}x{D} -> Stage-2 ranker on {K_TOP} candidates -> top-10 on the host",
               "unit": "ms per request (wall clock, host to host)"}
        for name, ranker_mod in (("stage1_only", None), ("two_stage", rk)):
            inf = AdRecommenderInference(model_dir="/nonexistent", device=str(dev), preprocessor=_Pre(),
                                         two_tower_model=tt, transformer_ranker=ranker_mod, faiss_index=flat_index,
                                         verbose=False)
            for u in users[:8]:
                inf.recommend_ads(u, top_k=10, stage1_k=K_TOP)
            lat, s1, s2 = [], [], []
            for r in range(200):
                t0 = time.perf_counter()
                rec = inf.recommend_ads(users[r % 64], top_k=10, stage1_k=K_TOP)
                lat.append((time.perf_counter() - t0) * 1e3)
                s1.append(rec["timing"]["stage1_ms"])
                s2.append(rec["timing"]["stage2_ms"])
            t0 = time.perf_counter()
            inf.batch_recommend(users, top_k=10, stage1_k=K_TOP)
            b64 = (time.perf_counter() - t0) * 1e3
            out[name] = {"median_ms": float(np.median(lat)), "p99_ms": float(np.percentile(lat, 99)),
                         "stage1_median_ms": float(np.median(s1)), "stage2_median_ms": float(np.median(s2)),
                         "batch_recommend_64_users_ms": b64, "requests_per_s_serial": 1e3 / float(np.median(lat))}
        out["value"] = out["two_stage"]["median_ms"]
        return out
    guarded("recommend_request", request)
    return extra


def run_anchor(args, torch, dev, peaks):
    """The N > 1 workload (100M x 256) on ONE GPU: fp16 scan copy 51 GB + fp32 master 102 GB of the 180 GB.
    The strong-scaling anchor the N = 2/4/8 lines are to be compared with."""
    from movie_recommender_demo_b200.faiss_retrieval import IndexFlatIP
    free, _ = torch.cuda.mem_get_info()
    need = CORPUS_MULTI * D * 6 + (8 << 30)
    if free < need:
        return {"skipped": f"needs {need / 1e9:.0f} GB of free device memory, {free / 1e9:.0f} GB available"}
    idx = IndexFlatIP(D, device=dev.index)
    idx.reserve(CORPUS_MULTI)
    gen = torch.Generator(device=dev)
    t0 = time.time()
    for c in range((CORPUS_MULTI + CHUNK - 1) // CHUNK):
        idx.add(_chunk_rows(torch, dev, gen, c)[: min(CHUNK, CORPUS_MULTI - c * CHUNK)], normalize=True)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    out = {"workload": f"Flat IP top-{K_TOP} over {CORPUS_MULTI}x{D} on ONE B200 (the N>1 workload unsharded)",
           "unit": "queries/s", "corpus_build_s": round(build_s, 1), "batches": {}}
    qg = torch.Generator(device=dev).manual_seed(2)
    for Q in (64, args.batch):
        qs = torch.randn((Q, D), generator=qg, device=dev)
        ms = _timed(torch, lambda: idx.search_device(qs, K_TOP, normalize=True), 2, 5)
        _, _, st, _ = idx.search_device(qs, K_TOP, normalize=True)
        out["batches"][str(Q)] = {"value": Q / ms * 1e3, "ms_per_step": ms, "queries_not_provably_exact": int((st != 0).sum())}
    out["value"] = out["batches"][str(args.batch)]["value"]
    out["ms_per_step"] = out["batches"][str(args.batch)]["ms_per_step"]
    return out


EXTRA_NAMES = ["flat_q1", "flat_q64", "flat_clustered", "ivf_10M", "ivfpq_10M", "tower_65536", "stage2_ranker",
               "recommend_request", "flat_100M_one_gpu"]


def run_extras_isolated(args, emit_err, timeout_s=900):
    """Run the sub-results in a child `bench.py --extras-child FILE`.  The child rewrites FILE after every
    sub-result; when it dies (device fault, OOM kill, timeout) the sub-result it was on is recorded as an error
    and a new child carries on with the rest."""
    import tempfile
    extra = {}
    skip = ["flat_100M_one_gpu"] if args.no_anchor else []
    fd, path = tempfile.mkstemp(prefix="b2r_extras_", suffix=".json")
    os.close(fd)
    try:
        for attempt in range(4):
            if all(n in skip for n in EXTRA_NAMES):
                break
            Path(path).write_text("{}")
            cmd = [sys.executable, str(ROOT / "bench.py"), "--extras-child", path, "--extras-skip", ",".join(skip),
                   "--batch", str(args.batch), "--scan-dtype", args.scan_dtype]
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
            try:
                r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=timeout_s, env=env)
                rc, err_tail = r.returncode, (r.stderr or "")[-400:]
            except subprocess.TimeoutExpired as exc:
                rc, err_tail = -9, f"timed out after {timeout_s} s: " + str(exc.stderr or "")[-300:]
            try:
                state = json.loads(Path(path).read_text() or "{}")
            except Exception:
                state = {}
            extra.update(state.get("done", {}))
            skip += [n for n in state.get("done", {}) if n not in skip]
            running = state.get("running")
            if rc == 0 and not running:
                break
            if running and running not in extra:
                extra[running] = {"error": f"the child process running this sub-result died (rc {rc}): {err_tail}"[:600]}
                emit_err(f"extra {running}: child died with rc {rc}; continuing with the rest in a new process")
                skip.append(running)
            elif not running:      # died before / between sub-results: do not loop on it
                emit_err(f"extras child failed with rc {rc}: {err_tail}")
                extra.setdefault("_child_error", f"rc {rc}: {err_tail}"[:600])
                break
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass
    return {n: extra[n] for n in EXTRA_NAMES if n in extra} | {k: v for k, v in extra.items() if k not in EXTRA_NAMES}


def run_extras_child(args):
    """Child side of run_extras_isolated: same 1M-row corpus as the headline, then every sub-result not skipped."""
    import torch
    from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
    os.dup2(2, 1)                      # native banners must not reach the parent's JSON stream
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    peaks = _peaks()
    FAISSIndex.verbose = False
    skip = [n for n in args.extras_skip.split(",") if n]
    path = Path(args.extras_child)
    state = {"done": {}, "running": None}

    def save():
        tmp = path.with_suffix(".tmp")
        tmp.write_text(json.dumps(state))
        os.replace(tmp, path)

    def progress(name, extra):
        state["running"] = name
        state["done"] = dict(extra)
        save()

    def emit_err(msg):
        sys.stderr.write(f"[bench extras] {msg}\n")

    index = FAISSIndex(D, 'Flat', device=0)
    index.index.reserve(CORPUS_1GPU)
    if args.scan_dtype != "auto":
        index.index.set_param("scan_dtype", {"bf16": 0, "fp16": 1}[args.scan_dtype])
    g = torch.Generator(device=dev)
    for c in range((CORPUS_1GPU + CHUNK - 1) // CHUNK):
        index.add(_chunk_rows(torch, dev, g, c)[: min(CHUNK, CORPUS_1GPU - c * CHUNK)])
    torch.cuda.synchronize()
    q_dev_extra = torch.randn((max(args.batch, 64), D), generator=torch.Generator().manual_seed(2)).to(dev)
    extra = run_extras(args, torch, dev, peaks, index, q_dev_extra, emit_err, skip=skip, progress=progress)
    del index
    torch.cuda.empty_cache()
    if "flat_100M_one_gpu" not in skip:
        progress("flat_100M_one_gpu", extra)
        try:
            extra["flat_100M_one_gpu"] = run_anchor(args, torch, dev, peaks)
        except Exception as exc:
            extra["flat_100M_one_gpu"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        torch.cuda.synchronize()
        progress(None, extra)
    return 0


# ------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from movie_recommender_demo_b200 import _lib
    from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex, _pipe_chunks
    from movie_recommender_demo_b200.sharded import ShardedFlatIndex, exchange_mode, shard_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything native libraries print to fd 1 meanwhile (NCCL's
    # "NCCL version ..." banner at NCCL_DEBUG=WARN/VERSION) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = _peaks()
    FAISSIndex.verbose = False

    def emit_err(msg):
        sys.stderr.write(f"[bench rank {rank}] {msg}\n")

    total_rows = args.corpus_rows or (CORPUS_1GPU if world == 1 else CORPUS_MULTI)
    lo_row, hi_row = shard_rows(total_rows, world, rank)
    Q = args.batch

    # ---- corpus shard: generated on the device chunk by chunk (seed = 100 + global chunk id), so the same
    #      rows exist for every world size; never materialised on the host
    t_build = time.time()
    if world == 1:
        index = FAISSIndex(D, 'Flat', device=local_rank)   # reference surface: default ids -> device id map
        flat = index.index
        flat.reserve(total_rows)
        sharded = None
    else:
        sharded = ShardedFlatIndex(D, total_rows, device=local_rank)
        flat = sharded.local
        index = None
    if args.scan_dtype != "auto":
        flat.set_param("scan_dtype", {"bf16": 0, "fp16": 1}[args.scan_dtype])
    g = torch.Generator(device=dev)
    for c in range(lo_row // CHUNK, (hi_row - 1) // CHUNK + 1):
        rows = _chunk_rows(torch, dev, g, c)
        a = max(lo_row, c * CHUNK) - c * CHUNK
        b = min(hi_row, (c + 1) * CHUNK) - c * CHUNK
        if world == 1:
            index.add(rows[a:b])
        else:
            sharded.add_local(rows[a:b], normalize=True)
        del rows
    torch.cuda.synchronize()
    t_build = time.time() - t_build

    # ---- queries: pinned host memory (e2e) and a resident device copy (value)
    gq = torch.Generator().manual_seed(2)
    q_host = torch.randn((max(Q, 64), D), generator=gq)[:Q].contiguous().pin_memory()
    q_np = q_host.numpy()
    q_dev = q_host.to(dev)
    q_dev_extra = torch.randn((max(Q, 64), D), generator=torch.Generator().manual_seed(2)).to(dev)

    def device_step(queries, eager=False):
        if world == 1:
            if Q <= 256 and not eager:   # launch-bound regime: the library's captured-graph path (static result buffers)
                return flat.search_device_static(queries, K_TOP, normalize=True)
            Dl, Il, st, _ = flat.search_device(queries, K_TOP, normalize=True)
            return Dl, Il, st
        if eager:
            Dm, Im, st, _ = sharded._search_eager(queries, K_TOP, True)
            return Dm, Im, st
        return sharded.search_device(queries, K_TOP, normalize=True)   # local scan + packed exchange + merge

    def e2e_step():
        if world == 1:
            return index.search(q_np, k=K_TOP)       # numpy in -> numpy (ids, distances) out
        Dm, Im = sharded.search(q_np, K_TOP)          # host queries -> host result on every rank, collective retry
        return Im, Dm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    windows = []
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("B2R_BENCH_NO_SAMPLER") else None

    # ---- warm-up (also exercises the e2e path once)
    for _ in range(max(args.warmup, 3)):
        _, _, st = device_step(q_dev)
    if not args.no_e2e:
        e2e_step()
    barrier()
    # status is OR-ed across shards by the exchange: this is the all-rank count
    status_bad = int((st != 0).sum().item())

    def launch_count():
        return int(lib.b2r_debug_launch_count()) + flat.replayed_launches + (sharded.replayed_launches if sharded else 0)

    # ---- timed: value (device resident)
    launches0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        device_step(q_dev)
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    value_window = windows[-1]
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = launch_count() - launches0
    value = args.steps * Q / (ms_total / 1e3)

    # ---- roofline: the filter-scan kernel alone, CUDA events around each launch on its stream
    flat.set_param("profile", args.steps)
    barrier()
    w0 = time.time()
    for _ in range(args.steps):
        device_step(q_dev, eager=True)    # per-launch CUDA events cannot be read back from a graph replay
    barrier()
    windows.append((w0, time.time()))
    scan_ms = flat.get_param("scan_ms_avg")
    scan_ms = max_over_ranks(scan_ms)
    flat.set_param("profile", 0)
    n_shard = hi_row - lo_row
    clocks_value = sampler.summary([value_window]) if sampler else None
    if Q >= 256:
        flops = 2.0 * Q * n_shard * D
        achieved = flops / (scan_ms / 1e3) / 1e12
        peak, peak_kind = _pick_tensor_peak(peaks, clocks_value, ms_total / 1e3)
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_kind": peak_kind, "frac_of_burst_peak": achieved / peaks["bf16_tflops"],
                "frac_of_sustained_peak": achieved / peaks["bf16_tflops_sustained"],
                "whole_step_frac_of_burst_peak": flops / (ms_total / args.steps / 1e3) / 1e12 / peaks["bf16_tflops"]}
    else:
        nbytes = n_shard * D * 2.0 + Q * D * 2.0
        achieved = nbytes / (scan_ms / 1e3) / 1e9
        peak = peaks["hbm_gbs"]
        step_bytes = n_shard * D * 2.0 + Q * (K_TOP * D * 4 + D * 4 + K_TOP * 12)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_kind": f"{peaks['source']} STREAM-style copy",
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_total / args.steps / 1e3) / 1e9,
                               "frac": step_bytes / (ms_total / args.steps / 1e3) / 1e9 / peak}}
        if achieved > peak:   # a read-only stream is not bound by the read+write copy figure
            roof["note"] = ("above 1.0: the kernel only READS the corpus; the measured peak is a copy (read + write) "
                            "bandwidth, which a pure read stream exceeds on long launches")
    traffic = _traffic_note(Q) if n_shard == CORPUS_1GPU else None
    pair = Q > 128 and int(flat.get_param("pair_scan")) == 1
    roof.update({"kernel": ("scan_pair_kernel (tcgen05 cta_group::2 score contraction on CTA pairs + threshold filter)" if pair
                            else "scan_tc_kernel<MQ,FILTER> (tcgen05 score contraction + threshold filter)"),
                 "kernel_ms": scan_ms, "traffic": traffic,
                 "traffic_source": ("dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this "
                                    "kernel at this batch (profiles/traffic.json); not measured in this run") if traffic else None})

    # ---- timed: e2e through the public API with host buffers
    e2e_value = e2e_ms = None
    res = None
    if not args.no_e2e:
        # warm-up in the same pattern as the timed loop (the previous result stays alive while the next one
        # is produced, so TWO sets of pinned result buffers must exist before the clock starts)
        for _ in range(max(args.warmup, 3)):
            res = e2e_step()
        barrier()
        w0 = time.time()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        windows.append((w0, time.time()))
        e2e_value = args.steps * Q / (e2e_ms / 1e3)
    e2e_retries = (flat.last_retries if world == 1 else sharded.last_retries) if not args.no_e2e else None

    clocks = sampler.stop(windows) if sampler else None

    # ---- CPU baseline + parity spot check (rank 0)
    cpu = None
    parity = None
    if rank == 0 and res is not None and not args.no_parity:
        try:
            if world == 1 and total_rows <= 4 * CORPUS_1GPU:
                from oracle.compare import compare_topk
                from oracle.flat import normalize_L2, topk_desc
                x_host = index.index.reconstruct_n(0, total_rows).cpu().numpy()
                # 64 queries spread over every chunk of the pipelined host-result path and every 128-query block
                sel = np.unique(np.minimum(np.arange(64) * max(Q // 64, 1) + (np.arange(64) * 7) % max(Q // 64, 1), Q - 1))
                qn = q_np[sel].astype(np.float32).copy()
                normalize_L2(qn)
                rd, rid = topk_desc(qn @ x_host.T, K_TOP, extra=32)
                r = compare_topk(res[0][sel], res[1][sel], rid, rd, K_TOP, gap_tol=1e-6)
                chunks = _pipe_chunks(Q) if Q >= 2048 else [(0, Q)]
                parity = (f"pass ({len(sel)} queries from {sum(1 for a, b in chunks if ((sel >= a) & (sel < b)).any())} of "
                          f"{len(chunks)} pipeline chunks vs the CPU oracle over the full corpus: ids/order identical outside "
                          f"1e-6 gaps, max |score err| {r['max_abs_score_err']:.1e})")
            else:
                parity = parity_by_regeneration(torch, dev, total_rows, q_np, res[0], res[1])
        except AssertionError as exc:  # report, never hide
            parity = f"FAIL: {exc}"
    if world == 1 and rank == 0 and not args.no_cpu and total_rows <= 4 * CORPUS_1GPU:
        x_host = index.index.reconstruct_n(0, total_rows).cpu().numpy()
        search = cpu_search_fn(x_host, index.id_map)
        search(q_np[:8], K_TOP)
        piece = min(Q, 512)
        n_done, t0 = 0, time.perf_counter()
        while True:
            lo = n_done % Q
            search(q_np[lo:lo + piece], K_TOP)
            n_done += min(piece, Q - lo)
            dt = time.perf_counter() - t0
            if dt > 12.0 or n_done >= 2 * Q:
                break
        cpu = {"value": n_done / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{piece}-query pieces of the batch vs the full {total_rows}x{D} corpus, "
                         f"{n_done} queries in {dt:.1f}s",
               "note": "oracle port (torch-CPU sgemm + topk + python id remap); faiss-cpu is not installable here"}
        del x_host, search

    # ---- the other BASELINE configs + the single-GPU anchor of the sharded workload (N = 1 only)
    extra = None
    scan_fp16 = int(flat.get_param("scan_dtype")) == 1
    if world == 1 and rank == 0 and not args.no_extra and total_rows == CORPUS_1GPU:
        # The sub-results run in CHILD processes: a device fault in one of them (a sticky CUDA error kills every
        # later call of its process) must not cost the headline measurement above, nor the other sub-results.
        if os.environ.get("B2R_BENCH_EXTRAS_INLINE"):    # A/B knob: the sub-results in THIS process
            extra = run_extras(args, torch, dev, peaks, index, q_dev_extra, emit_err)
            del index, flat
            torch.cuda.empty_cache()
            if not args.no_anchor:
                extra["flat_100M_one_gpu"] = run_anchor(args, torch, dev, peaks)
        else:
            del index, flat
            torch.cuda.empty_cache()
            extra = run_extras_isolated(args, emit_err)

    if rank == 0:
        exch = None if world == 1 else exchange_mode(Q, world)
        out = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f16" if scan_fp16 else "bf16", "data": "synthetic",
            "config": {
                "workload": (f"Flat IP top-{K_TOP} over {total_rows}x{D} ad corpus, query batch {Q}"
                             + (f", row-sharded over {world} B200, packed {exch} exchange + merge" if world > 1 else ", single B200")),
                "corpus_rows": total_rows, "rows_per_gpu": n_shard, "dim": D, "k": K_TOP, "batch": Q,
                "arithmetic": ("fp16" if scan_fp16 else "bf16")
                + " operands / fp32 accumulate tcgen05 scan (unit-norm rows), exact fp32 rescore of the final candidates",
                "l2_policy": "no flush: corpus (16-bit scan copy + fp32 master) is larger than the 126 MB L2",
                "corpus_build_s": round(t_build, 2),
                "scaling_note": ("the N=1 line is BASELINE configs[1] (1M-row corpus on one GPU); N>1 lines are configs[4] "
                                 "(the 100M-row corpus row-sharded over N GPUs, strong scaling).  value_N / (N * value_1) "
                                 "therefore compares different problems: use extra.flat_100M_one_gpu of the N=1 line (the "
                                 "same 100M-row workload on ONE GPU) as the anchor of the N>1 lines, or scan_throughput "
                                 "(query x rows / s), which is workload-independent"),
                "launch": ("CUDA-graph replay of the step's launch sequence (batch <= 256)" if Q <= 256 else "eager stream launches"),
            },
            "e2e": None if e2e_value is None else {"value": e2e_value, "unit": "queries/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * K_TOP * 12 + Q * 4,
                    "retries_in_last_call": e2e_retries},
            "gpu_launches": launches,
            # workload-independent rate (the N=1 and N>1 configs differ in corpus size): query x row scores per second
            "scan_throughput": {"value": value * total_rows, "unit": "query*rows/s"},
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "queries_not_provably_exact": status_bad,
            "queries_not_provably_exact_scope": "all ranks (status words are OR-ed across shards by the exchange)" if world > 1 else "this GPU",
            "parity_spot_check": parity,
            "extra": extra,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        # captured graphs hold NCCL kernels: release them before the communicator goes away, and never let a
        # teardown problem turn a finished measurement into a hung job
        sharded.release_graphs()
        dist.barrier()
        torch.cuda.synchronize()
        import threading
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(timeout=30)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def _traffic_note(Q):
    """dram bytes per launch of the filter-scan kernel from the committed ncu capture, if any."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(str(Q))
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096, help="queries per step")
    ap.add_argument("--corpus-rows", type=int, default=0, help="override total corpus rows")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer (e2e) leg: profiling runs only")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity spot check")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (N = 1)")
    ap.add_argument("--no-anchor", action="store_true", help="skip the 100M-rows-on-one-GPU anchor (N = 1)")
    ap.add_argument("--extras-child", default="", help=argparse.SUPPRESS)   # internal: see run_extras_isolated
    ap.add_argument("--extras-skip", default="", help=argparse.SUPPRESS)
    ap.add_argument("--scan-dtype", choices=["auto", "bf16", "fp16"], default="auto",
                    help="16-bit format of the scan copy (auto = fp16 for L2-normalised corpora)")
    args = ap.parse_args()
    if args.extras_child:
        return run_extras_child(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
