"""CPU tests of the Stage-2 ranker's test infrastructure and host logic:
the numpy oracle is PINNED against outputs of the reference's own module (tests/golden/ranker_*.npz,
made by tests/golden/make_ranker_golden.py), the B200 module keeps the reference's state-dict keys, and the
host-side weight fold (attention over one key == one linear map) is exact."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
from weights import RANKER_CONFIGS, feature_dims, make_ranker_state  # noqa: E402

TASKS = ("ctr", "engagement", "revenue")


def _golden(name):
    return np.load(ROOT / "tests" / "golden" / f"ranker_{name}.npz")


@pytest.mark.parametrize("name", list(RANKER_CONFIGS))
@pytest.mark.parametrize("tag,cross_std", [("", None), ("_rawinit", 1.0)])
def test_oracle_reproduces_the_reference_ranker(name, tag, cross_std):
    from oracle.ranker import ranker_forward
    cfg, gold = RANKER_CONFIGS[name], _golden(name)
    state = make_ranker_state(cfg, int(gold["seed"]), cross_std)
    out = ranker_forward(state, gold["ucat"], gold["acat"], gold["num"], cfg["num_heads"])
    for t in TASKS:
        ref = gold[t + tag]
        # fp32 both sides, different summation order (numpy vs ATen): ~1e-6 relative to the output scale
        assert np.abs(out[t] - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), t


def test_oracle_raises_on_out_of_range_ids():
    from oracle.ranker import ranker_forward
    cfg, gold = RANKER_CONFIGS["small"], _golden("small")
    state = make_ranker_state(cfg, int(gold["seed"]))
    bad = gold["acat"].copy()
    bad[3, 1] = cfg["ad_cards"][1]
    with pytest.raises(IndexError):
        ranker_forward(state, gold["ucat"], bad, gold["num"], cfg["num_heads"])


@pytest.mark.parametrize("name", list(RANKER_CONFIGS))
def test_module_keeps_reference_state_dict_keys_and_fold_is_exact(name):
    import torch
    from movie_recommender_demo_b200.transformer_ranker import TransformerRanker, fold_ranker_weights
    from oracle.ranker import embed_features, ranker_forward
    cfg, gold = RANKER_CONFIGS[name], _golden(name)
    user, ad = feature_dims(cfg)
    m = TransformerRanker(user, ad, cfg["numerical_dim"], embedding_dim=cfg["embedding_dim"], d_model=cfg["d_model"],
                          num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], d_ff=cfg["d_ff"])
    assert list(m.state_dict().keys()) == list(gold["keys"])          # names AND order of the reference module
    state = make_ranker_state(cfg, int(gold["seed"]))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    a = fold_ranker_weights(m)
    # evaluate the FOLDED arrays (what b2r_ranker_create receives) in float64 and compare with the oracle's
    # full attention arithmetic: pins the fold, the cross transpose and the head stacking
    x = embed_features(state, gold["ucat"], gold["acat"], gold["num"]).astype(np.float64)
    f = {k: v.astype(np.float64) for k, v in a.items()}
    ln = lambda v, g, b: (v - v.mean(1, keepdims=True)) / np.sqrt(v.var(1, keepdims=True) + 1e-5) * g + b   # noqa: E731
    x = x @ f["w_proj"].T + f["b_proj"]
    for l in range(cfg["num_layers"]):
        x = ln(x + x @ f["w_attn"][l].T + f["b_attn"][l], f["ln1_g"][l], f["ln1_b"][l])
        ff = np.maximum(x @ f["w_fc1"][l].T + f["b_fc1"][l], 0) @ f["w_fc2"][l].T + f["b_fc2"][l]
        x = ln(x + ff, f["ln2_g"][l], f["ln2_b"][l])
    x0 = xl = x
    for c in range(3):
        xl = x0 * (xl @ f["w_cross"][c].T + f["b_cross"][c]) + xl
    ref = ranker_forward(state, gold["ucat"], gold["acat"], gold["num"], cfg["num_heads"])
    for t, task in enumerate(TASKS):
        h = np.maximum(xl @ f["w_h1"][t].T + f["b_h1"][t], 0)
        h = np.maximum(h @ f["w_h2"][t].T + f["b_h2"][t], 0)
        got = h @ f["w_h3"][t] + f["b_h3"][t]
        assert np.abs(got - ref[task]).max() <= 2e-5 * max(1.0, np.abs(ref[task]).max()), task


def test_ranker_refuses_cpu_and_train_mode():
    import torch
    from movie_recommender_demo_b200.transformer_ranker import TransformerRanker
    m = TransformerRanker({"a": 4}, {"b": 4}, 2, embedding_dim=8, d_model=128, num_heads=4, num_layers=1, d_ff=128)
    z = torch.zeros((2, 1), dtype=torch.int64)
    with pytest.raises(RuntimeError, match="eval"):
        m(z, z, torch.zeros((2, 2)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.eval()(z, z, torch.zeros((2, 2)))
