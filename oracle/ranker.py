"""numpy fp32 restatement of the reference's Stage-2 TransformerRanker in eval mode (TEST INFRASTRUCTURE).

Follows transformer_ranker.py — the FULL attention arithmetic, not the folded form the B200 path uses, so that
the fold itself is under test:
  * embed_features            :312-330  user tables, ad tables (ModuleDict order), numericals, concatenated
  * feature_projection + pos  :352-361  x = W f + b ; x = x.unsqueeze(1) + positional_encoding[:, :1]
  * MultiHeadAttention        :56-90    Q/K/V projections, heads, softmax(QK^T / sqrt(d_k)), W_o  (seq_len 1)
  * TransformerEncoderLayer   :146-153  x = norm1(x + attn) ; x = norm2(x + fc2(relu(fc1 x)))   (LayerNorm eps 1e-5)
  * FeatureInteractionLayer   :199-207  xl = x0 * (xl @ W_i + b_i) + xl
  * prediction heads          :283-310, :372-376  Linear ReLU Linear ReLU Linear, squeeze(1)
Dropout is the identity in eval mode.
PINNED against tests/golden/ranker_*.npz, which hold outputs of the reference's own module
(tests/golden/make_ranker_golden.py).
"""
from __future__ import annotations

import re

import numpy as np

LN_EPS = np.float32(1e-5)


def _lin(state, prefix, x):
    return x @ np.asarray(state[prefix + ".weight"], np.float32).T + np.asarray(state[prefix + ".bias"], np.float32)


def _layer_norm(state, prefix, x):
    mean = x.mean(axis=-1, keepdims=True, dtype=np.float32)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True, dtype=np.float32)
    return (x - mean) / np.sqrt(var + LN_EPS) * np.asarray(state[prefix + ".weight"], np.float32) + \
        np.asarray(state[prefix + ".bias"], np.float32)


def embed_features(state, ucat, acat, num):
    cols = []
    for group, cat in (("user_embeddings", ucat), ("ad_embeddings", acat)):
        pat = re.compile(re.escape(group) + r"\.(.+)\.weight$")
        tables = [state[k] for k in state if pat.match(k)]
        for f, W in enumerate(tables):
            idx = cat[:, f].astype(np.int64)
            if (idx < 0).any() or (idx >= W.shape[0]).any():
                raise IndexError("index out of range in self")
            cols.append(np.asarray(W, np.float32)[idx])
    cols.append(num.astype(np.float32))
    return np.concatenate(cols, axis=1)


def attention(state, prefix, x, num_heads):
    """x [B, 1, d]; the reference's multi-head attention written out for a general sequence length"""
    B, S, d = x.shape
    dk = d // num_heads
    split = lambda t: t.reshape(B, S, num_heads, dk).transpose(0, 2, 1, 3)   # noqa: E731
    Q, K, V = (split(_lin(state, f"{prefix}.{n}", x)) for n in ("W_q", "W_k", "W_v"))
    scores = Q @ K.transpose(0, 1, 3, 2) / np.float32(np.sqrt(dk))
    scores = scores - scores.max(axis=-1, keepdims=True)
    w = np.exp(scores)
    w = w / w.sum(axis=-1, keepdims=True)
    ctx = (w @ V).transpose(0, 2, 1, 3).reshape(B, S, d)
    return _lin(state, f"{prefix}.W_o", ctx)


def ranker_forward(state: dict, ucat, acat, num, num_heads: int = 8) -> dict:
    x = _lin(state, "feature_projection", embed_features(state, ucat, acat, num))
    x = x[:, None, :] + np.asarray(state["positional_encoding"], np.float32)[:, :1, :]
    layers = sorted({int(m.group(1)) for k in state for m in [re.match(r"transformer_layers\.(\d+)\.", k)] if m})
    for l in layers:
        p = f"transformer_layers.{l}"
        x = _layer_norm(state, f"{p}.norm1", x + attention(state, f"{p}.self_attention", x, num_heads))
        ff = _lin(state, f"{p}.feed_forward.fc2", np.maximum(_lin(state, f"{p}.feed_forward.fc1", x), np.float32(0)))
        x = _layer_norm(state, f"{p}.norm2", x + ff)
    x = x[:, 0, :]
    x0, xl = x, x
    crosses = sorted({int(m.group(1)) for k in state
                      for m in [re.match(r"feature_interaction\.cross_weights\.(\d+)$", k)] if m})
    for i in crosses:
        W = np.asarray(state[f"feature_interaction.cross_weights.{i}"], np.float32)
        b = np.asarray(state[f"feature_interaction.cross_biases.{i}"], np.float32)
        xl = x0 * (xl @ W + b) + xl
    out = {}
    tasks = list(dict.fromkeys(m.group(1) for k in state for m in [re.match(r"prediction_heads\.([^.]+)\.", k)] if m))
    for t in tasks:
        h = np.maximum(_lin(state, f"prediction_heads.{t}.0", xl), np.float32(0))
        h = np.maximum(_lin(state, f"prediction_heads.{t}.3", h), np.float32(0))
        out[t] = _lin(state, f"prediction_heads.{t}.6", h)[:, 0]
    return out
