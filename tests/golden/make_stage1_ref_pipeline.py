"""Generates tests/golden/stage1_ref_pipeline.npz: Stage 1 of the train.py-shaped system run ENTIRELY through the
reference's own code in this container - `two_tower_model.TwoTowerModel` (towers), `faiss_retrieval.FAISSIndex`
(corpus build: `add(ad_emb, ad_ids)`) and `faiss_retrieval.TwoStageRetriever.retrieve_and_rank` (one user per
call, no ad-feature lookup -> the stage-1 list is returned, faiss_retrieval.py:302-327) - with only faiss's
IndexFlatIP / normalize_L2 replaced by the numpy stand-in of make_wrapper_golden.py (faiss is not installable here).

Same seeds, sizes and ids as make_stage1_golden.py, so tests/test_oracle_cpu.py can check that the fixture the B200
path is compared with (stage1_cfg1.npz, produced by reference towers + the oracle wrapper) is what the reference
pipeline returns.  Run in the build container only.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_stage1_golden import AD_SEED, K, N_ADS, N_USERS, STATE_SEED, USER_SEED  # noqa: E402
from make_wrapper_golden import _install_faiss_stand_in  # noqa: E402
from weights import CONFIGS, feature_dims, make_inputs, make_state  # noqa: E402


def main():
    _install_faiss_stand_in()
    sys.path.insert(0, "/root/reference")
    import faiss_retrieval as ref_fr      # the reference, untouched
    import two_tower_model as ref_tt

    cfg = CONFIGS["cfg1"]
    user, ad = feature_dims(cfg)
    model = ref_tt.TwoTowerModel(user_feature_dims=user, ad_feature_dims=ad, numerical_dim=cfg["numerical_dim"],
                                 embedding_dim=cfg["embedding_dim"], hidden_dims=cfg["hidden_dims"],
                                 output_dim=cfg["output_dim"])
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in make_state(cfg, STATE_SEED).items()})
    model.eval()
    _, _, acat = make_inputs(cfg, AD_SEED, N_ADS)
    ucat, unum, _ = make_inputs(cfg, USER_SEED, N_USERS)
    # corpus build as training_pipeline.build_faiss_index does it (:515-539): tower batches -> vstack -> add
    embs = []
    with torch.no_grad():
        for lo in range(0, N_ADS, 1024):
            embs.append(model.get_ad_embeddings(torch.from_numpy(acat[lo:lo + 1024])).cpu().numpy())
    index = ref_fr.FAISSIndex(dimension=cfg["output_dim"], index_type='Flat')
    index.add(np.vstack(embs), [10 * i + 3 for i in range(N_ADS)])
    retriever = ref_fr.TwoStageRetriever(model, torch.nn.Identity(), index, device='cpu')
    ids, dist = [], []
    for u in range(N_USERS):
        a, s = retriever.retrieve_and_rank(torch.from_numpy(ucat[u:u + 1]), torch.from_numpy(unum[u:u + 1]), stage1_k=K)
        ids.append(a)
        dist.append(s)
    np.savez_compressed(HERE / "stage1_ref_pipeline.npz", ids=np.array(ids, dtype=np.int64),
                        dist=np.array(dist, dtype=np.float32), k=K, n_users=N_USERS, torch_version=torch.__version__)
    print("reference pipeline:", np.array(ids).shape)


if __name__ == "__main__":
    main()
