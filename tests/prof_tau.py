"""Experiment: filter-scan kernel time vs candidate rate (tau=+inf -> pure fast path)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from movie_recommender_demo_b200.faiss_retrieval import IndexFlatIP

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = 1_000_000
g = torch.Generator(device="cuda").manual_seed(1)
idx = IndexFlatIP(256)
idx.add(torch.randn((N, 256), generator=g, device="cuda"), normalize=True)
q = torch.randn((Q, 256), generator=g, device="cuda")
for tau_val, label in [(float("inf"), "no hits"), (0.26, "~40/query"), (0.215, "~300"), (0.2, "~700"), (0.185, "~1500"), (0.17, "~3200")]:
    tau = torch.full((Q,), tau_val, device="cuda")
    for _ in range(2):
        idx.search_device(q, 500, normalize=True, tau=tau)
    torch.cuda.synchronize()
    idx.set_param("profile", 5)
    for _ in range(5):
        D, I, st, tr = idx.search_device(q, 500, normalize=True, tau=tau)
    torch.cuda.synchronize()
    print(f"Q={Q} tau={tau_val} ({label}): filter scan {idx.get_param('scan_ms_avg'):.3f} ms")
