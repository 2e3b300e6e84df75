"""Tiny driver for profiling one Flat configuration: python tests/prof_flat.py Q [N] [steps]."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from movie_recommender_demo_b200.faiss_retrieval import IndexFlatIP

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(1)
idx = IndexFlatIP(256)
for lo in range(0, N, 1 << 20):
    n = min(1 << 20, N - lo)
    idx.add(torch.randn((n, 256), generator=g, device="cuda"), normalize=True)
q = torch.randn((Q, 256), generator=g, device="cuda")
for _ in range(2):
    idx.search_device(q, 500, normalize=True)
torch.cuda.synchronize()
idx.set_param("profile", steps)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    D, I, st, tr = idx.search_device(q, 500, normalize=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"Q={Q} N={N}: {ms:.3f} ms/step, filter scan kernel {idx.get_param('scan_ms_avg'):.3f} ms, "
      f"status_nonzero={(st != 0).sum().item()}")
