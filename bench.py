#!/usr/bin/env python
"""bench.py — Stage-1 retrieval throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--batch Q] [--impl reference]

metric  : retrieval queries/sec @ top-500 (Flat inner product, d=256)
N = 1   : BASELINE config[1] — 1M x 256 ad corpus on one B200, query batch Q (default 4096)
N > 1   : BASELINE config[4] — 100M x 256 corpus row-sharded over N ranks (strong scaling),
          every rank searches its shard for the same Q queries, per-rank top-500 merged by an
          NCCL all-gather + b2r_topk_merge on every rank.
A "step" = one pass of the hot path over one batch of Q queries:
  value : queries already resident in HBM, results left in HBM (CUDA-event timed, max over ranks)
  e2e   : FAISSIndex.search(numpy queries in pinned host memory) -> numpy ids + distances,
          host<->device copies and the id remap inside the timed region
  roofline    : the dominant kernel (filter scan, tcgen05) timed with CUDA events on its stream
  cpu_baseline: the reference-style CPU path (oracle port: fp32 sgemm + top-k + python id remap)
                on this box's host cores, on a bounded sample of the same workload
`--impl reference` times only that CPU path (the reference's faiss-cpu wheel is not installable
here: no network, no wheel — see DESIGN.md), same config / metric / unit.
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

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
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
    # bounded sample of the workload: a 256-query slice of the batch against (at most) a 1M-row slice of
    # the corpus per step; throughput is scaled to the whole corpus (a flat scan is linear in rows)
    n = min(total_rows, CORPUS_1GPU)
    scale = n / total_rows
    sample_q = min(args.batch, 256)
    x = make_host_corpus(n)
    search = cpu_search_fn(x, list(range(n)))
    rng = np.random.default_rng(2)
    q = rng.standard_normal((sample_q, D), dtype=np.float32)
    for _ in range(max(1, min(args.warmup, 2))):
        search(q, K_TOP)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        search(q, K_TOP)
    dt = time.perf_counter() - t0
    val = args.steps * sample_q / dt * scale
    cores = torch.get_num_threads()
    sample = (f"{sample_q} of {args.batch} queries per step vs a {n}-row slice of the {total_rows}x{D} corpus, "
              f"top-{K_TOP}" + (f"; queries/s scaled by {scale:g} to the full corpus" if scale != 1 else ""))
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"Flat IP top-{K_TOP} over {total_rows}x{D} ad corpus, query batch {args.batch}"
                                + (f", row-sharded over {world} B200 + NCCL all-gather merge" if world > 1
                                   else ", single B200") + " [CPU reference arm]"),
                   "corpus_rows": total_rows, "dim": D, "k": K_TOP, "batch": args.batch},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle port (torch-CPU sgemm + topk + python id remap); faiss-cpu is not installable here"},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from movie_recommender_demo_b200 import _lib
    from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
    from movie_recommender_demo_b200.sharded import ShardedFlatIndex, shard_rows

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
        g.manual_seed(100 + c)
        rows = torch.randn((CHUNK, D), generator=g, device=dev)
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
    q_host = torch.randn((Q, D), generator=gq).pin_memory()
    q_np = q_host.numpy()
    q_dev = q_host.to(dev)
    stream_ptr = lambda: int(torch.cuda.current_stream(dev).cuda_stream)  # noqa: E731

    def device_step(queries, eager=False):
        if world == 1:
            if Q <= 256 and not eager:   # launch-bound regime: the library's captured-graph path (static result buffers)
                return flat.search_device_static(queries, K_TOP, normalize=True)
            Dl, Il, st, _ = flat.search_device(queries, K_TOP, normalize=True)
            return Dl, Il, st
        return sharded.search_device(queries, K_TOP, normalize=True)   # local scan + all-gather + merge

    def e2e_step():
        if world == 1:
            return index.search(q_np, k=K_TOP)       # numpy in -> numpy (ids, distances) out
        qd = q_host.to(dev, non_blocking=True)
        Dm, Im, _ = device_step(qd)
        if rank == 0:
            return Im.cpu().numpy(), Dm.cpu().numpy()
        return None

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
    status_bad = int((st != 0).sum().item())

    # ---- timed: value (device resident)
    launches0 = int(lib.b2r_debug_launch_count()) + flat.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        device_step(q_dev)
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib.b2r_debug_launch_count()) + flat.replayed_launches - launches0
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
    shard_rows = hi_row - lo_row
    if Q >= 256:
        flops = 2.0 * Q * shard_rows * D
        achieved = flops / (scan_ms / 1e3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_kind": f"{peaks['source']} cuBLAS bf16, sustained (kernel timed inside a {args.steps}-step loop)",
                "frac_of_burst_peak": achieved / peaks["bf16_tflops"]}
    else:
        nbytes = shard_rows * D * 2.0 + Q * D * 2.0
        achieved = nbytes / (scan_ms / 1e3) / 1e9
        peak = peaks["hbm_gbs"]
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_kind": f"{peaks['source']} STREAM-style copy"}
        if achieved > peak:   # a read-only stream is not bound by the read+write copy figure
            roof["note"] = ("above 1.0: the kernel only READS the corpus; the measured peak is a copy (read + write) "
                            "bandwidth, which a pure read stream exceeds on long launches")
    roof.update({"kernel": "scan_tc_kernel<MQ,FILTER> (tcgen05 score contraction + threshold filter)",
                 "kernel_ms": scan_ms, "traffic": _traffic_note(Q) if shard_rows == CORPUS_1GPU else None})

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

    clocks = sampler.stop(windows) if sampler else None

    # ---- CPU baseline + parity spot check (rank 0, single GPU only)
    cpu = None
    parity = None
    if world == 1 and rank == 0 and not args.no_cpu and res is not None:
        from oracle.compare import compare_topk
        x_host = index.index.reconstruct_n(0, total_rows).cpu().numpy()
        search = cpu_search_fn(x_host, index.id_map)
        sample_q = min(Q, 256)
        search(q_np[:8], K_TOP)
        n_done, t0 = 0, time.perf_counter()
        while True:
            ids_c, d_c = search(q_np[:sample_q], K_TOP)
            n_done += sample_q
            dt = time.perf_counter() - t0
            if dt > 10.0 or n_done >= 16 * sample_q:
                break
        cpu = {"value": n_done / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{sample_q}-query slices of the batch vs the full {total_rows}x{D} corpus, "
                         f"{n_done} queries in {dt:.1f}s",
               "note": "oracle port (torch-CPU sgemm + topk + python id remap); faiss-cpu is not installable here"}
        try:
            # oracle with 32 extra ranks so near-ties at the boundary are comparable
            S = torch.from_numpy(q_np[:8] / np.linalg.norm(q_np[:8], axis=1, keepdims=True)) @ torch.from_numpy(x_host).T
            Dv, Iv = torch.topk(S, K_TOP + 32, dim=1)
            compare_topk(res[0][:8], res[1][:8], Iv.numpy(), Dv.numpy(), K_TOP, gap_tol=1e-6)
            parity = "pass (8 queries vs CPU oracle: ids/order identical outside 1e-6 gaps)"
        except AssertionError as exc:  # report, never hide
            parity = f"FAIL: {exc}"

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f16" if int(flat.get_param("scan_dtype")) == 1 else "bf16", "data": "synthetic",
            "config": {
                "workload": (f"Flat IP top-{K_TOP} over {total_rows}x{D} ad corpus, query batch {Q}"
                             + (f", row-sharded over {world} B200 + NCCL all-gather merge" if world > 1 else ", single B200")),
                "corpus_rows": total_rows, "rows_per_gpu": shard_rows, "dim": D, "k": K_TOP, "batch": Q,
                "arithmetic": ("fp16" if int(flat.get_param("scan_dtype")) == 1 else "bf16")
                + " operands / fp32 accumulate tcgen05 scan (unit-norm rows), exact fp32 rescore of the final candidates",
                "l2_policy": "no flush: corpus (bf16 scan copy + fp32 master) is larger than the 126 MB L2",
                "corpus_build_s": round(t_build, 2),
                "scaling_note": ("the N=1 line is BASELINE configs[1] (1M-row corpus on one GPU); N>1 lines are "
                                 "configs[4] (the 100M-row corpus row-sharded over N GPUs, strong scaling): compare "
                                 "the N>1 lines among themselves, or scan_throughput (query*rows/s) across all N"),
                "launch": ("CUDA-graph replay of the search's launch sequence (batch <= 256)" if world == 1 and Q <= 256
                           else "eager stream launches"),
            },
            "e2e": None if e2e_value is None else {"value": e2e_value, "unit": "queries/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * K_TOP * 12 + Q * 4},
            "gpu_launches": launches,
            # workload-independent rate (the N=1 and N>1 configs differ in corpus size): query x row scores per second
            "scan_throughput": {"value": value * total_rows, "unit": "query*rows/s"},
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "queries_not_provably_exact": status_bad,
            "parity_spot_check": parity,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
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
    ap.add_argument("--scan-dtype", choices=["auto", "bf16", "fp16"], default="auto",
                    help="16-bit format of the scan copy (auto = fp16 for L2-normalised corpora)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
