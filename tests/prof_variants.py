"""A/B driver for tunables of the flat search: one corpus, several parameter sets, per-set step time,
filter-scan kernel time and the number of queries the kernels flagged (status != 0).

    python tests/prof_variants.py Q N steps  "epi_warps=8"  "epi_warps=16"  "epi_warps=16,cand_factor=2.0"
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from movie_recommender_demo_b200.faiss_retrieval import IndexFlatIP

Q, N, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
variants = sys.argv[4:] or [""]
g = torch.Generator(device="cuda").manual_seed(1)
idx = IndexFlatIP(256)
for lo in range(0, N, 1 << 20):
    n = min(1 << 20, N - lo)
    idx.add(torch.randn((n, 256), generator=g, device="cuda"), normalize=True)
q = torch.randn((Q, 256), generator=g, device="cuda")
ref = None
for rep in range(2):                      # two rounds: order effects (clocks, power cap) show up as a spread
    for v in variants:
        for kv in filter(None, v.split(",")):
            name, val = kv.split("=")
            idx.set_param(name, float(val))
        for _ in range(3):
            D, I, st, tr = idx.search_device(q, 500, normalize=True)
        torch.cuda.synchronize()
        idx.set_param("profile", steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            D, I, st, tr = idx.search_device(q, 500, normalize=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        scan = idx.get_param("scan_ms_avg")
        idx.set_param("profile", 0)
        if ref is None:
            ref = (D.clone(), I.clone())
        same = bool((I == ref[1]).all().item()) and bool((D == ref[0]).all().item())
        print(f"[{v or 'default':40s}] Q={Q} N={N}: {ms:.3f} ms/step  scan {scan:.3f} ms  rest {ms - scan:.3f} ms  "
              f"flagged={(st != 0).sum().item()}  identical_to_first={same}", flush=True)
