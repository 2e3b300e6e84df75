"""numpy fp32 restatement of the reference towers in eval mode (TEST INFRASTRUCTURE).

Follows two_tower_model.py:
  * EmbeddingLayer.forward  :42-47   out[:, f*E:(f+1)*E] = W_f[cat[:, f]] in ModuleDict order, concat
  * UserTower.forward       :110-121 x = [emb ‖ num] -> mlp -> F.normalize(p=2, dim=1)
  * AdTower.forward         :175-184 same without numericals
  * mlp                     :83-95   Linear, BatchNorm1d (eval: running stats, eps=1e-5), ReLU,
                                     Dropout (eval: identity), ..., Linear
PINNED against tests/golden/towers_*.npz, which hold outputs of the reference's own code
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import re

import numpy as np

BN_EPS = np.float32(1e-5)
NORM_EPS = np.float32(1e-12)


def embedding_concat(state: dict, prefix: str, cat: np.ndarray) -> np.ndarray:
    """state keys `<prefix>.embedding_layer.embeddings.<name>.weight`, in insertion order."""
    pat = re.compile(re.escape(prefix) + r"\.embedding_layer\.embeddings\.(.+)\.weight$")
    tables = [state[k] for k in state if pat.match(k)]
    cols = []
    for f, W in enumerate(tables):
        idx = cat[:, f].astype(np.int64)
        if (idx < 0).any() or (idx >= W.shape[0]).any():
            raise IndexError("index out of range in self")
        cols.append(np.asarray(W, dtype=np.float32)[idx])
    return np.concatenate(cols, axis=1)


def mlp_forward(state: dict, prefix: str, x: np.ndarray) -> np.ndarray:
    pos = sorted({int(m.group(1)) for k in state
                  for m in [re.match(re.escape(prefix) + r"\.mlp\.(\d+)\.weight$", k)] if m})
    x = x.astype(np.float32)
    for p in pos:
        W = np.asarray(state[f"{prefix}.mlp.{p}.weight"], dtype=np.float32)
        if W.ndim == 2:  # Linear
            x = x @ W.T + np.asarray(state[f"{prefix}.mlp.{p}.bias"], dtype=np.float32)
        else:            # BatchNorm1d (eval) followed by ReLU
            mean = np.asarray(state[f"{prefix}.mlp.{p}.running_mean"], dtype=np.float32)
            var = np.asarray(state[f"{prefix}.mlp.{p}.running_var"], dtype=np.float32)
            beta = np.asarray(state[f"{prefix}.mlp.{p}.bias"], dtype=np.float32)
            x = (x - mean) / np.sqrt(var + BN_EPS) * W + beta
            x = np.maximum(x, np.float32(0))
    return x


def l2_normalize(x: np.ndarray) -> np.ndarray:
    n = np.sqrt(np.einsum("ij,ij->i", x, x, dtype=np.float32))
    return x / np.maximum(n, NORM_EPS)[:, None]


def tower_forward(state: dict, prefix: str, cat: np.ndarray, num: np.ndarray | None = None) -> np.ndarray:
    x = embedding_concat(state, prefix, cat)
    if num is not None:
        x = np.concatenate([x, num.astype(np.float32)], axis=1)
    return l2_normalize(mlp_forward(state, prefix, x))
