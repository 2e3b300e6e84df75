"""Generates tests/golden/towers_<cfg>.npz by running the UNMODIFIED reference towers
(`/root/reference/two_tower_model.py`) on deterministic weights and inputs.

Run in the build container only (`python tests/golden/make_golden.py`); the fixtures travel,
the reference does not.  Stored per config: state-dict key list (pins names and order),
inputs, the reference EmbeddingLayer output (bit-exact contract) and the reference tower
outputs (fp32, eval mode).
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, "/root/reference")
import two_tower_model as ref  # noqa: E402  (the reference, untouched)
from weights import CONFIGS, feature_dims, make_inputs, make_state  # noqa: E402

SEED = 20240
BATCH = 48


def main():
    only = sys.argv[1:]        # `python make_golden.py deep3 one_hidden` regenerates just those fixtures
    for name, cfg in CONFIGS.items():
        if only and name not in only:
            continue
        user, ad = feature_dims(cfg)
        model = ref.TwoTowerModel(user_feature_dims=user, ad_feature_dims=ad, numerical_dim=cfg["numerical_dim"],
                                  embedding_dim=cfg["embedding_dim"], hidden_dims=cfg["hidden_dims"],
                                  output_dim=cfg["output_dim"])
        state = make_state(cfg, SEED)
        keys = list(model.state_dict().keys())
        assert sorted(keys) == sorted(state.keys()), set(keys) ^ set(state.keys())
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()})
        model.eval()
        ucat, unum, acat = make_inputs(cfg, SEED, BATCH)
        with torch.no_grad():
            u_emb_layer = model.user_tower.embedding_layer(torch.from_numpy(ucat)).numpy()
            a_emb_layer = model.ad_tower.embedding_layer(torch.from_numpy(acat)).numpy()
            u = model.get_user_embeddings(torch.from_numpy(ucat), torch.from_numpy(unum)).numpy()
            a = model.get_ad_embeddings(torch.from_numpy(acat)).numpy()
        np.savez_compressed(HERE / f"towers_{name}.npz", seed=SEED, keys=np.array(keys), ucat=ucat, unum=unum,
                            acat=acat, user_embedding_layer=u_emb_layer, ad_embedding_layer=a_emb_layer,
                            user_out=u, ad_out=a, torch_version=torch.__version__)
        print(name, "user", u.shape, "ad", a.shape, "norms", np.linalg.norm(u, axis=1)[:3])


if __name__ == "__main__":
    main()
