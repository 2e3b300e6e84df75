"""Deterministic tower weights for the golden fixtures (numpy PCG64, stable across versions).

`make_state(cfg, seed)` returns {state_dict key: np.ndarray} for a TwoTowerModel with the
reference's key names (two_tower_model.py: `user_tower.embedding_layer.embeddings.<f>.weight`,
`user_tower.mlp.{0,1,4,5,8}.*`, same for `ad_tower`).  Used by make_golden.py (loaded into the
REFERENCE model to produce expected outputs) and by the tests (loaded into the oracle and into
the B200 modules), so the fixture only has to store inputs and expected outputs.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # train.py / inference.py shape: 6 user fields, 20 ad fields, 13 numericals
    "cfg1": dict(user_cards=[1000, 500, 100, 50, 1000, 500],
                 ad_cards=[100, 50, 1000, 500] * 4 + [100, 50, 20, 10],
                 numerical_dim=13, embedding_dim=16, hidden_dims=[512, 256], output_dim=256),
    # tutorial.ipynb shape
    "small": dict(user_cards=[30, 20, 10, 5, 30, 7], ad_cards=[11, 7, 5, 3] * 5,
                  numerical_dim=13, embedding_dim=16, hidden_dims=[256, 128], output_dim=128),
}


def feature_dims(cfg):
    user = {f"C{i + 1}": c for i, c in enumerate(cfg["user_cards"])}
    ad = {f"C{i + 1 + len(cfg['user_cards'])}": c for i, c in enumerate(cfg["ad_cards"])}
    return user, ad


def _tower_state(prefix, dims, in_extra, cfg, rng, out):
    E = cfg["embedding_dim"]
    for name, card in dims.items():
        out[f"{prefix}.embedding_layer.embeddings.{name}.weight"] = rng.standard_normal((card, E)).astype(np.float32)
    width = len(dims) * E + in_extra
    pos = 0
    for h in cfg["hidden_dims"]:
        bound = 1.0 / np.sqrt(width)
        out[f"{prefix}.mlp.{pos}.weight"] = rng.uniform(-bound, bound, (h, width)).astype(np.float32)
        out[f"{prefix}.mlp.{pos}.bias"] = rng.uniform(-bound, bound, h).astype(np.float32)
        out[f"{prefix}.mlp.{pos + 1}.weight"] = rng.uniform(0.5, 1.5, h).astype(np.float32)
        out[f"{prefix}.mlp.{pos + 1}.bias"] = (0.1 * rng.standard_normal(h)).astype(np.float32)
        out[f"{prefix}.mlp.{pos + 1}.running_mean"] = (0.1 * rng.standard_normal(h)).astype(np.float32)
        out[f"{prefix}.mlp.{pos + 1}.running_var"] = rng.uniform(0.5, 1.5, h).astype(np.float32)
        out[f"{prefix}.mlp.{pos + 1}.num_batches_tracked"] = np.array(7, dtype=np.int64)
        width = h
        pos += 4
    bound = 1.0 / np.sqrt(width)
    out[f"{prefix}.mlp.{pos}.weight"] = rng.uniform(-bound, bound, (cfg["output_dim"], width)).astype(np.float32)
    out[f"{prefix}.mlp.{pos}.bias"] = rng.uniform(-bound, bound, cfg["output_dim"]).astype(np.float32)


def make_state(cfg, seed: int):
    rng = np.random.default_rng(seed)
    user, ad = feature_dims(cfg)
    out = {}
    _tower_state("user_tower", user, cfg["numerical_dim"], cfg, rng, out)
    _tower_state("ad_tower", ad, 0, cfg, rng, out)
    return out


def make_inputs(cfg, seed: int, batch: int):
    rng = np.random.default_rng(seed + 1000)
    ucat = np.stack([rng.integers(0, c, batch) for c in cfg["user_cards"]], axis=1).astype(np.int64)
    acat = np.stack([rng.integers(0, c, batch) for c in cfg["ad_cards"]], axis=1).astype(np.int64)
    unum = rng.standard_normal((batch, cfg["numerical_dim"])).astype(np.float32)
    return ucat, unum, acat
