"""Deterministic tower weights for the golden fixtures (numpy PCG64, stable across versions).

`make_state(cfg, seed)` returns {state_dict key: np.ndarray} for a TwoTowerModel with the
reference's key names (two_tower_model.py: `user_tower.embedding_layer.embeddings.<f>.weight`,
`user_tower.mlp.{0,1,4,5,8}.*` for two hidden layers, same for `ad_tower`).  Used by make_golden.py (loaded into the
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
    # `hidden_dims` of other lengths (train.py:350 --hidden_dims nargs='+'; two_tower_model.py:83-95 loops over it)
    "deep3": dict(user_cards=[30, 20, 10, 5], ad_cards=[11, 7, 5, 3, 9],
                  numerical_dim=13, embedding_dim=16, hidden_dims=[384, 200, 128], output_dim=64),
    "one_hidden": dict(user_cards=[30, 20, 10, 5], ad_cards=[11, 7, 5, 3, 9],
                       numerical_dim=13, embedding_dim=16, hidden_dims=[192], output_dim=96),
    "no_hidden": dict(user_cards=[30, 20, 10, 5], ad_cards=[11, 7, 5, 3, 9],
                      numerical_dim=13, embedding_dim=8, hidden_dims=[], output_dim=32),
    # two hidden layers too wide for the fused kernel (layer-by-layer launches)
    "wide2": dict(user_cards=[30, 20, 10, 5], ad_cards=[11, 7, 5, 3, 9],
                  numerical_dim=13, embedding_dim=16, hidden_dims=[640, 384], output_dim=256),
    # widths that are not multiples of anything convenient (output rows not 16-byte aligned)
    "odd_out": dict(user_cards=[30, 20, 10, 5], ad_cards=[11, 7, 5, 3, 9],
                    numerical_dim=13, embedding_dim=16, hidden_dims=[100, 60], output_dim=50),
}
DEPTH_CONFIGS = ["deep3", "one_hidden", "no_hidden", "wide2", "odd_out"]


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


# ---- Stage-2 ranker (transformer_ranker.py) -----------------------------------------------------------------
RANKER_CONFIGS = {
    # inference.py:120-131 shape: embedding_dim 32, d_model 256, 8 heads, 3 layers, d_ff 1024
    "cfg1": dict(user_cards=[1000, 500, 100, 50, 1000, 500], ad_cards=[200] * 20, numerical_dim=13,
                 embedding_dim=32, d_model=256, num_heads=8, num_layers=3, d_ff=1024),
    "small": dict(user_cards=[30, 20, 10], ad_cards=[11, 7, 5, 3], numerical_dim=5,
                  embedding_dim=8, d_model=128, num_heads=4, num_layers=2, d_ff=256),
}


def make_ranker_state(cfg, seed: int, cross_std=None):
    """{state_dict key: np.ndarray} with the reference's key names.  Linear layers use torch's default bound
    1/sqrt(fan_in); the cross weights are N(0, cross_std^2) — the reference initialises them with randn
    (std 1, transformer_ranker.py:180-187), which blows activations up by ~16x per cross layer; the default
    here (1/sqrt(d)) stands for a trained checkpoint, `cross_std=1.0` reproduces the raw initialisation."""
    rng = np.random.default_rng(seed)
    d, E, dff = cfg["d_model"], cfg["embedding_dim"], cfg["d_ff"]
    out = {}

    def lin(name, n_out, n_in):
        bound = 1.0 / np.sqrt(n_in)
        out[name + ".weight"] = rng.uniform(-bound, bound, (n_out, n_in)).astype(np.float32)
        out[name + ".bias"] = rng.uniform(-bound, bound, n_out).astype(np.float32)

    out["positional_encoding"] = rng.standard_normal((1, 50, d)).astype(np.float32)
    user, ad = feature_dims(cfg)
    for n, c in user.items():
        out[f"user_embeddings.{n}.weight"] = rng.standard_normal((c, E)).astype(np.float32)
    for n, c in ad.items():
        out[f"ad_embeddings.{n}.weight"] = rng.standard_normal((c, E)).astype(np.float32)
    lin("feature_projection", d, (len(user) + len(ad)) * E + cfg["numerical_dim"])
    for l in range(cfg["num_layers"]):
        p = f"transformer_layers.{l}"
        for n in ("W_q", "W_k", "W_v", "W_o"):
            lin(f"{p}.self_attention.{n}", d, d)
        lin(f"{p}.feed_forward.fc1", dff, d)
        lin(f"{p}.feed_forward.fc2", d, dff)
        for n in ("norm1", "norm2"):
            out[f"{p}.{n}.weight"] = rng.uniform(0.5, 1.5, d).astype(np.float32)
            out[f"{p}.{n}.bias"] = (0.1 * rng.standard_normal(d)).astype(np.float32)
    cs = (1.0 / np.sqrt(d)) if cross_std is None else cross_std
    for i in range(3):
        out[f"feature_interaction.cross_weights.{i}"] = (cs * rng.standard_normal((d, d))).astype(np.float32)
    for i in range(3):
        out[f"feature_interaction.cross_biases.{i}"] = (cs * rng.standard_normal(d)).astype(np.float32)
    for t in ("ctr", "engagement", "revenue"):
        lin(f"prediction_heads.{t}.0", 256, d)
        lin(f"prediction_heads.{t}.3", 64, 256)
        lin(f"prediction_heads.{t}.6", 1, 64)
    return out


def make_ranker_inputs(cfg, seed: int, batch: int):
    rng = np.random.default_rng(seed + 2000)
    ucat = np.stack([rng.integers(0, c, batch) for c in cfg["user_cards"]], axis=1).astype(np.int64)
    acat = np.stack([rng.integers(0, c, batch) for c in cfg["ad_cards"]], axis=1).astype(np.int64)
    num = rng.standard_normal((batch, cfg["numerical_dim"])).astype(np.float32)
    return ucat, acat, num
