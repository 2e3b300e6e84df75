"""Generates tests/golden/ranker_<cfg>.npz by running the UNMODIFIED reference Stage-2 ranker
(`/root/reference/transformer_ranker.py`) in eval mode on deterministic weights and inputs.

Run in the build container only (`python tests/golden/make_ranker_golden.py`); the fixtures travel, the
reference does not.  Stored per config: state-dict key list (pins names and order), inputs, the reference's
raw head outputs for the "trained-like" cross weights and for the reference's raw randn initialisation.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, "/root/reference")
import transformer_ranker as ref  # noqa: E402  (the reference, untouched)
from weights import RANKER_CONFIGS, feature_dims, make_ranker_inputs, make_ranker_state  # noqa: E402

SEED = 31337
BATCH = 500          # stage1_k candidates of one user (inference.py:240-255)


def main():
    for name, cfg in RANKER_CONFIGS.items():
        user, ad = feature_dims(cfg)
        model = ref.TransformerRanker(user, ad, cfg["numerical_dim"], embedding_dim=cfg["embedding_dim"],
                                      d_model=cfg["d_model"], num_heads=cfg["num_heads"], num_layers=cfg["num_layers"],
                                      d_ff=cfg["d_ff"])
        keys = list(model.state_dict().keys())
        ucat, acat, num = make_ranker_inputs(cfg, SEED, BATCH)
        saved = dict(seed=SEED, keys=np.array(keys), ucat=ucat, acat=acat, num=num, torch_version=torch.__version__)
        for tag, cross_std in (("", None), ("_rawinit", 1.0)):
            state = make_ranker_state(cfg, SEED, cross_std)
            assert sorted(keys) == sorted(state.keys()), set(keys) ^ set(state.keys())
            model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()})
            model.eval()
            with torch.no_grad():
                pred = model(torch.from_numpy(ucat), torch.from_numpy(acat), torch.from_numpy(num))
            for t, v in pred.items():
                saved[f"{t}{tag}"] = v.numpy()
            print(name, tag or "trained-like", {t: (float(v.abs().max()), float(v.std())) for t, v in pred.items()})
        np.savez_compressed(HERE / f"ranker_{name}.npz", **saved)


if __name__ == "__main__":
    main()
