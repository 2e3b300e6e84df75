"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/b2retr.h declares; host-side module surfaces match the reference's names."""
import ctypes
import inspect

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(built_lib):
    from movie_recommender_demo_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 25
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in b2retr.h but not exported: {missing}"
    assert built_lib.b2r_version() == 100


def test_library_contains_blackwell_instructions(built_lib):
    """SASS must show tcgen05.mma (UTCHMMA), TMA (UTMALDG) and TMEM loads (LDTM)."""
    import shutil
    import subprocess
    from movie_recommender_demo_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass


def test_errors_are_returned_not_thrown(built_lib):
    from movie_recommender_demo_b200 import _lib
    assert built_lib.b2r_index_search(None, 1, None, 0, 10, 0, None, None, None, None, None, None, 0, None) == _lib.EINVAL
    assert b"NULL handle" in built_lib.b2r_last_error()
    assert built_lib.b2r_topk_merge(0, 1, 1, None, None, None, None, 1, None) == _lib.EINVAL


def test_faiss_index_surface_matches_reference():
    from movie_recommender_demo_b200 import faiss_retrieval as fr
    sig = inspect.signature(fr.FAISSIndex.__init__)
    assert list(sig.parameters)[:6] == ["self", "dimension", "index_type", "nlist", "nprobe", "use_gpu"]
    assert sig.parameters["index_type"].default == 'IVF' and sig.parameters["nlist"].default == 100
    assert sig.parameters["nprobe"].default == 10 and sig.parameters["use_gpu"].default is False
    for name in ("train", "add", "search", "batch_search", "save", "load", "get_stats"):
        assert callable(getattr(fr.FAISSIndex, name))
    s = inspect.signature(fr.FAISSIndex.search)
    assert s.parameters["k"].default == 100 and s.parameters["return_distances"].default is True
    assert inspect.signature(fr.FAISSIndex.batch_search).parameters["batch_size"].default == 1000
    r = inspect.signature(fr.TwoStageRetriever.retrieve_and_rank)
    assert r.parameters["stage1_k"].default == 500 and r.parameters["stage2_k"].default == 10


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FAISSIndex(64, 'Flat')
    with pytest.raises(ValueError, match="Unknown index type: Bogus"):
        FAISSIndex(64, 'Bogus')
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FAISSIndex(64, 'HNSW')          # exact-L2 stand-in on the same kernels: needs the GPU like the rest


def test_tower_modules_keep_reference_state_dict_keys():
    """Same keys/shapes as the reference model (pinned in the golden fixture's key list)."""
    import torch
    from movie_recommender_demo_b200.two_tower_model import TwoTowerModel
    from weights import CONFIGS, feature_dims, make_state
    golden = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "towers_cfg1.npz")
    cfg = CONFIGS["cfg1"]
    user, ad = feature_dims(cfg)
    m = TwoTowerModel(user, ad, cfg["numerical_dim"], cfg["embedding_dim"], cfg["hidden_dims"], cfg["output_dim"])
    assert list(m.state_dict().keys()) == golden["keys"].tolist()
    state = make_state(cfg, int(golden["seed"]))
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()})   # strict
    assert m.output_dim == 256 and m.user_tower.output_dim == 256
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.get_ad_embeddings(torch.zeros(2, 20, dtype=torch.long))
    m.train()
    with pytest.raises(RuntimeError, match="eval-mode"):
        m.get_ad_embeddings(torch.zeros(2, 20, dtype=torch.long))


def test_bn_folding_matches_oracle_math():
    import torch
    from movie_recommender_demo_b200.two_tower_model import TwoTowerModel, fold_tower_weights
    from oracle import towers as otowers
    from weights import CONFIGS, feature_dims, make_inputs, make_state
    cfg = CONFIGS["small"]
    user, ad = feature_dims(cfg)
    m = TwoTowerModel(user, ad, cfg["numerical_dim"], cfg["embedding_dim"], cfg["hidden_dims"], cfg["output_dim"])
    state = make_state(cfg, 3)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()})
    ucat, unum, _ = make_inputs(cfg, 3, 16)
    x = np.concatenate([otowers.embedding_concat(state, "user_tower", ucat), unum], axis=1)
    ref = otowers.mlp_forward(state, "user_tower", x)
    (w1, b1), (w2, b2), (w3, b3) = fold_tower_weights(m.user_tower.mlp)
    h = np.maximum(x @ w1.T + b1, 0)
    h = np.maximum(h @ w2.T + b2, 0)
    got = h @ w3.T + b3
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)
