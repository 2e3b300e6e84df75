import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from movie_recommender_demo_b200.faiss_retrieval import IndexFlatIP
g = torch.Generator(device="cuda").manual_seed(1)
idx = IndexFlatIP(256)
idx.add(torch.randn((1_000_000, 256), generator=g, device="cuda"), normalize=True)
for Q in (64, 256, 384, 512, 768, 1024, 2048):
    q = torch.randn((Q, 256), generator=g, device="cuda")
    for _ in range(3):
        idx.search_device(q, 500, normalize=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        idx.search_device(q, 500, normalize=True)
    t_issue = (time.perf_counter() - t0) / 20 * 1e3
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / 20 * 1e3
    print(f"Q={Q}: cpu issue {t_issue:.3f} ms/call, total {t_total:.3f} ms/call, ws={idx._ws.numel()/1e6:.0f} MB")
