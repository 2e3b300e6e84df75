"""Generates tests/golden/stage1_cfg1.npz: the Stage-1 answer (top-k ad ids + scores per user) for the
train.py-shaped model, with the embeddings produced by the UNMODIFIED reference towers
(`/root/reference/two_tower_model.py`, imported here) and the retrieval step by the order of operations of
the reference wrapper (faiss_retrieval.py:97-166: astype float32 -> normalize_L2 -> IndexFlatIP add/search ->
id_map remap).  faiss itself is not installable in this image, so that last step is the numpy restatement in
oracle/flat.py - the fixture pins the towers and the end-to-end glue, not faiss's arithmetic.

Run in the build container only (`python tests/golden/make_stage1_golden.py`); the fixture travels, the
reference does not.  Stored: seeds, sizes, the reference tower outputs for the users, and k + EXTRA results
per user (the extra ranks let the comparator resolve near-ties at the k boundary).
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")
import two_tower_model as ref  # noqa: E402  (the reference, untouched)
from oracle.flat import OracleFAISSIndex  # noqa: E402
from weights import CONFIGS, feature_dims, make_inputs, make_state  # noqa: E402

STATE_SEED, AD_SEED, USER_SEED = 4101, 4102, 4103
N_ADS, N_USERS, K, EXTRA = 6000, 24, 100, 24


def main():
    cfg = CONFIGS["cfg1"]
    user, ad = feature_dims(cfg)
    model = ref.TwoTowerModel(user_feature_dims=user, ad_feature_dims=ad, numerical_dim=cfg["numerical_dim"],
                              embedding_dim=cfg["embedding_dim"], hidden_dims=cfg["hidden_dims"],
                              output_dim=cfg["output_dim"])
    state = make_state(cfg, STATE_SEED)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()})
    model.eval()
    _, _, acat = make_inputs(cfg, AD_SEED, N_ADS)
    ucat, unum, _ = make_inputs(cfg, USER_SEED, N_USERS)
    with torch.no_grad():
        ad_emb = model.get_ad_embeddings(torch.from_numpy(acat)).numpy()
        user_emb = model.get_user_embeddings(torch.from_numpy(ucat), torch.from_numpy(unum)).numpy()
    ad_ids = [10 * i + 3 for i in range(N_ADS)]                 # non-trivial id_map
    index = OracleFAISSIndex(cfg["output_dim"], 'Flat')
    index.add(ad_emb, ad_ids)
    ids, dist = index.search(user_emb, k=K, extra=EXTRA)
    np.savez_compressed(HERE / "stage1_cfg1.npz", state_seed=STATE_SEED, ad_seed=AD_SEED, user_seed=USER_SEED,
                        n_ads=N_ADS, n_users=N_USERS, k=K, extra=EXTRA, user_out=user_emb,
                        ids=ids.astype(np.int64), dist=dist.astype(np.float32), torch_version=torch.__version__)
    gaps = -np.diff(dist[:, :K + 1], axis=1)
    print("stage1 golden:", ids.shape, "min adjacent gap", gaps.min(), "median", np.median(gaps))


if __name__ == "__main__":
    main()
