"""Profiling driver for the IVF paths: python tests/prof_ivf.py KIND Q [N] [nlist] [steps]"""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
FAISSIndex.verbose = False
kind = sys.argv[1] if len(sys.argv) > 1 else "IVF"
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000
nlist = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
sample = int(sys.argv[6]) if len(sys.argv) > 6 else 1
debug = int(sys.argv[7]) if len(sys.argv) > 7 else 0
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
centres = torch.randn((nlist, 256), generator=g, device=dev)
x = torch.empty((N, 256), device=dev)
for lo in range(0, N, 1 << 20):
    hi = min(N, lo + (1 << 20))
    x[lo:hi] = centres[torch.randint(0, nlist, (hi - lo,), generator=g, device=dev)] + 0.35 * torch.randn((hi - lo, 256), generator=g, device=dev)
x = torch.nn.functional.normalize(x, dim=1)
idx = FAISSIndex(256, kind, nlist=nlist, nprobe=32, pq_m=32)
idx.index.nprobe = 32
idx.add(x)
idx.index.set_param('ivf_sample', sample)
idx.index.set_param('ivf_debug', debug)
idx.index.set_param('profile', 0)
import os
fused = int(os.environ.get('B2R_IVF_FUSED', '1'))
if kind == 'IVF':
    idx.index.set_param('ivf_fused', fused)
    idx.index.set_param('ivf_sample_rows', int(os.environ.get('B2R_IVF_SAMPLE_ROWS', '128')))
q = torch.nn.functional.normalize(centres[torch.randint(0, nlist, (Q,), generator=g, device=dev)] + 0.35 * torch.randn((Q, 256), generator=g, device=dev), dim=1)
for _ in range(2):
    idx.index.search_device(q, 500, normalize=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    idx.index.search_device(q, 500, normalize=True)
torch.cuda.synchronize()
ms_step = (time.perf_counter() - t0) / steps * 1e3        # BEFORE anything else touches the device
_, _, _st, _ = idx.index.search_device(q, 500, normalize=True)
print(f"flagged={int((_st != 0).sum())} sample_rows={os.environ.get('B2R_IVF_SAMPLE_ROWS', '128')}", end=" ")
print(f"fused={fused} debug={debug} {kind} N={N} nlist={nlist} Q={Q} sample={sample}: {ms_step:.3f} ms/step")
import os
if os.environ.get("PROF_TABLE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            idx.index.search_device(q, 500, normalize=True)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]
    for e in rows:
        print(f"  {e.key[:70]:70s} n={e.count:4d} total={e.device_time_total / steps / 1e3:8.3f} ms/step")
    evs = [e for e in prof.events() if e.device_time_total > 0 and "Memset" not in e.name and "cudaLaunch" not in e.name]
    per = len(evs) // steps
    print("  -- launch sequence of the last step (us):")
    for e in evs[-per:]:
        print(f"     {e.device_time_total:9.1f}  {e.name[:90]}")
