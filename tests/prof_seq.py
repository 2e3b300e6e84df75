"""Per-launch device times of one search step (torch.profiler / CUPTI, no ncu needed):
   python tests/prof_seq.py KIND N Q K [nlist] [nprobe] [pq_m]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile
from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
FAISSIndex.verbose = False
kind = sys.argv[1]
N, Q, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
nlist = int(sys.argv[5]) if len(sys.argv) > 5 else 1024
nprobe = int(sys.argv[6]) if len(sys.argv) > 6 else 32
pq_m = int(sys.argv[7]) if len(sys.argv) > 7 else 32
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
idx = FAISSIndex(256, kind, nlist=nlist, nprobe=nprobe, pq_m=pq_m)
if kind == "Flat":
    for lo in range(0, N, 1 << 20):
        idx.add(torch.randn((min(N, lo + (1 << 20)) - lo, 256), generator=g, device=dev))
    q = torch.randn((Q, 256), generator=g, device=dev)
else:
    centres = torch.randn((nlist, 256), generator=g, device=dev)
    x = torch.empty((N, 256), device=dev)
    for lo in range(0, N, 1 << 20):
        hi = min(N, lo + (1 << 20))
        x[lo:hi] = centres[torch.randint(0, nlist, (hi - lo,), generator=g, device=dev)] + 0.35 * torch.randn((hi - lo, 256), generator=g, device=dev)
    idx.add(torch.nn.functional.normalize(x, dim=1))
    q = centres[torch.randint(0, nlist, (Q,), generator=g, device=dev)] + 0.35 * torch.randn((Q, 256), generator=g, device=dev)
    idx.index.nprobe = nprobe
for _ in range(3):
    idx.index.search_device(q, K, normalize=True)
torch.cuda.synchronize()
steps = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    idx.index.search_device(q, K, normalize=True)
e1.record()
torch.cuda.synchronize()
print(f"{kind} N={N} Q={Q} k={K}: {e0.elapsed_time(e1) / steps:.3f} ms/step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        idx.index.search_device(q, K, normalize=True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time_total > 0 and "Memset" not in e.name and "cudaLaunch" not in e.name]
per = len(evs) // steps
for e in evs[-per:]:
    print(f"   {e.device_time_total:9.1f} us  {e.name[:100]}")
