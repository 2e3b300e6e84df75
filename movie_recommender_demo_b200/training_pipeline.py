"""Drop-in for `training_pipeline.build_faiss_index` (training_pipeline.py:488-546), the
corpus-build caller of the hot path.

Same signature and result (an `'IVF'`, nlist=100, nprobe=10 `FAISSIndex` holding one embedding
per dataset row, ids = row numbers, saved to `save_path`), but device-resident: the AdTower
output of every batch goes straight from the tower kernels into `FAISSIndex.add` as a CUDA
tensor — no `.cpu().numpy()` per batch, no `np.vstack`, no host faiss.  The trainers of the
reference module are out of scope and are not provided here.
"""
from __future__ import annotations

import torch

from .faiss_retrieval import FAISSIndex

__all__ = ["build_faiss_index"]


def build_faiss_index(model, ad_data, device: str, save_path: str = None, batch_size: int = 1024,
                      index_type: str = 'IVF', nlist: int = 100, nprobe: int = 10) -> FAISSIndex:
    """`ad_data`: a torch Dataset whose items are dicts with key 'ad_categorical' (the reference's
    AdDataset), or an integer tensor / array [N, F] of ad categorical features."""
    say = print if FAISSIndex.verbose else (lambda *a, **k: None)
    say("\n=== Building FAISS Index ===")
    model = model.to(device).eval()
    if isinstance(ad_data, torch.utils.data.Dataset):
        loader = torch.utils.data.DataLoader(ad_data, batch_size=batch_size, shuffle=False)
        batches = (b['ad_categorical'] for b in loader)
        total = len(ad_data)
    else:
        feats = torch.as_tensor(ad_data)
        batches = (feats[i:i + batch_size] for i in range(0, len(feats), batch_size))
        total = len(feats)
    d = model.output_dim
    index = FAISSIndex(dimension=d, index_type=index_type, nlist=nlist, nprobe=nprobe)
    done = 0
    with torch.no_grad():
        if index_type == 'Flat':
            # a flat index needs no training: every batch goes from the tower kernels straight into the index's
            # pre-sized corpus buffers (fp32 master + 16-bit scan copy) - no intermediate embedding matrix at all
            index.index.reserve(total)
            for ad_cat in batches:
                emb = model.get_ad_embeddings(ad_cat.to(device))
                index.add(emb, list(range(done, done + len(emb))))
                done += len(emb)
        else:
            # IVF types train on the embeddings of the first add (faiss_retrieval.py:107-108): the reference hands
            # ALL rows to that one call, so they are collected first - in ONE pre-sized device tensor (a list of
            # chunks + torch.cat would hold the corpus twice)
            ad_embeddings = torch.empty((total, d), dtype=torch.float32, device=device)
            for ad_cat in batches:
                emb = model.get_ad_embeddings(ad_cat.to(device))
                ad_embeddings[done:done + len(emb)].copy_(emb)
                done += len(emb)
            index.add(ad_embeddings[:done], list(range(done)))
            del ad_embeddings
    say(f"Generated {done} ad embeddings")
    if save_path:
        index.save(save_path)
        say(f"✓ FAISS index saved to {save_path}")
    return index
