"""Phase timeline of the fused tower kernel (debug hook `trace_ptr`): clock64 stamps of the MMA issuer, the first
gather warp and the first epilogue warp for the first 4 CTAs.   python tests/prof_tower.py [--small]"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from movie_recommender_demo_b200 import _lib
from movie_recommender_demo_b200.two_tower_model import UserTower

rows = 1_000_000 if "--small" in sys.argv else 10_000_000
dev = torch.device("cuda")
torch.manual_seed(5)
t = UserTower({f"C{i + 1}": rows for i in range(26)}, 13).to(dev).eval()
if "--pair" in sys.argv:
    t.pair = 1
B = 65536
g = torch.Generator(device=dev).manual_seed(6)
cat = torch.randint(0, rows, (B, 26), generator=g, device=dev)
num = torch.randn((B, 13), generator=g, device=dev)
for _ in range(3):
    t(cat, num)
torch.cuda.synchronize()
trace = torch.zeros((4, 8, 64), dtype=torch.int64, device=dev)
lib = _lib.load()
_lib.check(lib.b2r_tower_set_param(t._handle, b"trace_ptr", float(trace.data_ptr())))
t(cat, num)
torch.cuda.synchronize()
_lib.check(lib.b2r_tower_set_param(t._handle, b"trace_ptr", 0.0))
tr = trace.cpu().numpy()
names = {0: "mma:tile_start", 1: "mma:tmem_free", 10: "mma:gemm1_issued", 11: "mma:h1[0..3]_ready", 12: "mma:gemm2_issued",
         13: "mma:gemm3_issued", 32: "epi:acc1_full", 33: "epi:epi1_done", 34: "epi:acc2_full", 35: "epi:epi2_done",
         36: "epi:acc3_full", 37: "epi:pass1_done", 38: "epi:pass2_done", 39: "epi:tile_done"}
for kc in range(8):
    names[2 + kc] = f"mma:a_full[{kc}]"
    names[16 + kc] = f"gat:chunk{kc}_start"
    names[24 + kc] = f"gat:chunk{kc}_arrive"
for cta in range(2):
    t0 = min(int(v) for v in tr[cta].reshape(-1) if v)
    for it in range(5):
        ev = [(int(v) - t0, names.get(s, str(s))) for s, v in enumerate(tr[cta, it]) if v]
        if not ev:
            continue
        print(f"--- CTA {cta} tile {it}")
        for c, n in sorted(ev):
            print(f"{c:9d}  {n}")
