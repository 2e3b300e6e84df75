"""Where does the end-to-end (host buffers) time go?  python tests/prof_e2e.py [Q]"""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
FAISSIndex.verbose = False
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
idx = FAISSIndex(256, 'Flat')
idx.add(torch.randn((1_000_000, 256), generator=g, device=dev))
qh = torch.randn((Q, 256)).pin_memory()
qn = qh.numpy()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
qd = qh.to(dev)
print("h2d queries            ", t(lambda: qh.to(dev)))
print("device search          ", t(lambda: idx.index.search_device(qd, 500, normalize=True)))
D, I, st, tr = idx.index.search_device(qd, 500, normalize=True)
def d2h():
    a = torch.empty(D.shape, dtype=D.dtype, pin_memory=True); a.copy_(D, non_blocking=True)
    b = torch.empty(I.shape, dtype=I.dtype, pin_memory=True); b.copy_(I, non_blocking=True)
    torch.cuda.synchronize()
print("d2h D+I (fresh pinned) ", t(d2h))
a = torch.empty(D.shape, dtype=D.dtype, pin_memory=True); b = torch.empty(I.shape, dtype=I.dtype, pin_memory=True)
def d2h2():
    a.copy_(D, non_blocking=True); b.copy_(I, non_blocking=True); torch.cuda.synchronize()
print("d2h D+I (reused pinned)", t(d2h2))
print("index.index.search(np) ", t(lambda: idx.index.search(qn, 500, normalize=True)))
print("FAISSIndex.search(np)  ", t(lambda: idx.search(qn, k=500)))

import os
for sizes in sys.argv[2:]:                       # explicit pipeline chunk sizes, e.g. 3584,512
    os.environ["B2R_PIPE_SIZES"] = sizes
    print(f"FAISSIndex.search(np) chunks {sizes:16s}", t(lambda: idx.search(qn, k=500)), flush=True)
os.environ.pop("B2R_PIPE_SIZES", None)
