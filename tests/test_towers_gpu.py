"""GPU parity tests of the tower path against the golden outputs of the reference's own
two_tower_model.py (tests/golden/towers_*.npz) and the numpy oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = __import__("pathlib").Path(__file__).parent / "golden"
TOWER_ATOL = 1e-3    # north_star: bf16 operands / fp32 accumulate, outputs are unit vectors (|x_i| <= 1)


def _model(cfg_name, seed=None):
    import torch
    from movie_recommender_demo_b200.two_tower_model import TwoTowerModel
    from weights import CONFIGS, feature_dims, make_state
    fx = np.load(GOLDEN / f"towers_{cfg_name}.npz")
    cfg = CONFIGS[cfg_name]
    user, ad = feature_dims(cfg)
    m = TwoTowerModel(user, ad, cfg["numerical_dim"], cfg["embedding_dim"], cfg["hidden_dims"], cfg["output_dim"])
    state = make_state(cfg, int(fx["seed"]) if seed is None else seed)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()})
    return m.to("cuda").eval(), fx, cfg, state


@pytest.mark.parametrize("cfg_name", ["cfg1", "small"])
def test_embedding_gather_is_bit_exact(built_lib, cfg_name):
    import torch
    m, fx, cfg, _ = _model(cfg_name)
    u = m.user_tower.embedding_layer(torch.from_numpy(fx["ucat"]).cuda()).cpu().numpy()
    a = m.ad_tower.embedding_layer(torch.from_numpy(fx["acat"]).cuda()).cpu().numpy()
    assert np.array_equal(u, fx["user_embedding_layer"])
    assert np.array_equal(a, fx["ad_embedding_layer"])


@pytest.mark.parametrize("cfg_name", ["cfg1", "small"])
def test_towers_match_reference_golden(built_lib, cfg_name):
    import torch
    m, fx, cfg, _ = _model(cfg_name)
    with torch.no_grad():
        u = m.get_user_embeddings(torch.from_numpy(fx["ucat"]).cuda(), torch.from_numpy(fx["unum"]).cuda())
        a = m.get_ad_embeddings(torch.from_numpy(fx["acat"]).cuda())
        u2, a2 = m(torch.from_numpy(fx["ucat"]).cuda(), torch.from_numpy(fx["unum"]).cuda(),
                   torch.from_numpy(fx["acat"]).cuda())
    u, a = u.cpu().numpy(), a.cpu().numpy()
    assert np.abs(u - fx["user_out"]).max() < TOWER_ATOL, np.abs(u - fx["user_out"]).max()
    assert np.abs(a - fx["ad_out"]).max() < TOWER_ATOL, np.abs(a - fx["ad_out"]).max()
    np.testing.assert_allclose(np.linalg.norm(u, axis=1), 1.0, atol=1e-5)
    assert np.array_equal(u2.cpu().numpy(), u) and np.array_equal(a2.cpu().numpy(), a)


@pytest.mark.parametrize("cfg_name", ["deep3", "one_hidden", "no_hidden", "wide2", "odd_out"])
def test_hidden_dims_of_any_length_match_reference_golden(built_lib, cfg_name):
    """`hidden_dims` with 0, 1, 3 entries, two entries too wide for the fused kernel, and widths that are not
    multiples of 4 (reference two_tower_model.py:83-95 builds one block per entry; train.py:350 exposes it):
    gather + one tcgen05 GEMM launch per Linear, against outputs of the reference's own module."""
    import torch
    from oracle import towers as otowers
    from weights import make_inputs
    m, fx, cfg, state = _model(cfg_name)
    ucat, unum, acat = (torch.from_numpy(fx[k]).cuda() for k in ("ucat", "unum", "acat"))
    with torch.no_grad():
        assert np.array_equal(m.user_tower.embedding_layer(ucat).cpu().numpy(), fx["user_embedding_layer"])
        u = m.get_user_embeddings(ucat, unum).cpu().numpy()
        a = m.get_ad_embeddings(acat).cpu().numpy()
    assert m.user_tower._handle is not None
    from movie_recommender_demo_b200 import _lib
    assert _lib.load().b2r_tower_get_param(m.user_tower._handle, b"fused") == 0.0
    assert u.shape == fx["user_out"].shape and a.shape == fx["ad_out"].shape
    assert np.abs(u - fx["user_out"]).max() < TOWER_ATOL, np.abs(u - fx["user_out"]).max()
    assert np.abs(a - fx["ad_out"]).max() < TOWER_ATOL, np.abs(a - fx["ad_out"]).max()
    # a batch that is not a multiple of the 128-row tile, against the oracle (pinned to the same goldens)
    ucat2, unum2, acat2 = make_inputs(cfg, 91, 333)
    with torch.no_grad():
        u2 = m.get_user_embeddings(torch.from_numpy(ucat2).cuda(), torch.from_numpy(unum2).cuda()).cpu().numpy()
        a2 = m.get_ad_embeddings(torch.from_numpy(acat2).cuda()).cpu().numpy()
    assert np.abs(u2 - otowers.tower_forward(state, "user_tower", ucat2, unum2)).max() < TOWER_ATOL
    assert np.abs(a2 - otowers.tower_forward(state, "ad_tower", acat2)).max() < TOWER_ATOL
    # the bf16 operand format on the same towers (what a saturating checkpoint switches to): 8x coarser rounding
    m.user_tower.operand_dtype = "bf16"
    m.user_tower._free()
    with torch.no_grad():
        ub = m.get_user_embeddings(ucat, unum).cpu().numpy()
    assert np.abs(ub - fx["user_out"]).max() < 2e-2


@pytest.mark.parametrize("B", [1, 2, 127, 128, 129, 1000, 4099])
def test_tower_batch_sizes_vs_oracle(built_lib, B):
    import torch
    from oracle import towers as otowers
    from weights import make_inputs
    m, fx, cfg, state = _model("cfg1")
    ucat, unum, acat = make_inputs(cfg, 77, B)
    with torch.no_grad():
        u = m.get_user_embeddings(torch.from_numpy(ucat).cuda(), torch.from_numpy(unum).cuda()).cpu().numpy()
        a = m.get_ad_embeddings(torch.from_numpy(acat).cuda()).cpu().numpy()
    assert np.abs(u - otowers.tower_forward(state, "user_tower", ucat, unum)).max() < TOWER_ATOL
    assert np.abs(a - otowers.tower_forward(state, "ad_tower", acat)).max() < TOWER_ATOL


def test_out_of_range_index_raises_like_torch(built_lib):
    import torch
    m, fx, cfg, _ = _model("small")
    bad = fx["ucat"].copy()
    bad[3, 2] = cfg["user_cards"][2]
    with pytest.raises(IndexError):
        m.user_tower.embedding_layer(torch.from_numpy(bad).cuda())
    with pytest.raises(IndexError):
        m.get_user_embeddings(torch.from_numpy(bad).cuda(), torch.from_numpy(fx["unum"]).cuda())


def test_reload_weights_rebuilds_native_tower(built_lib):
    import torch
    from weights import make_state
    m, fx, cfg, _ = _model("small")
    cat = torch.from_numpy(fx["acat"]).cuda()
    a1 = m.get_ad_embeddings(cat).cpu().numpy()
    state2 = make_state(cfg, 999)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state2.items()})
    a2 = m.get_ad_embeddings(cat).cpu().numpy()
    from oracle import towers as otowers
    assert np.abs(a2 - otowers.tower_forward(state2, "ad_tower", fx["acat"])).max() < TOWER_ATOL
    assert np.abs(a1 - a2).max() > 1e-2


def test_config4_shape_gather_26_fields_large_tables(built_lib):
    """BASELINE config 4 shape, scaled tables: 26 fields x 1M-row tables, 13 dense, batch 65536:
    gather bit-exact vs torch indexing; tower output unit-norm and equal to the oracle on a sample."""
    import torch
    from movie_recommender_demo_b200.two_tower_model import UserTower
    from oracle import towers as otowers
    torch.manual_seed(5)
    dims = {f"C{i+1}": 1_000_000 for i in range(26)}
    t = UserTower(dims, 13).to("cuda").eval()
    g = torch.Generator(device="cuda").manual_seed(6)
    B = 65536
    cat = torch.randint(0, 1_000_000, (B, 26), generator=g, device="cuda")
    num = torch.randn((B, 13), generator=g, device="cuda")
    emb = t.embedding_layer(cat)
    ref = torch.cat([e.weight[cat[:, i]] for i, e in enumerate(t.embedding_layer.embeddings.values())], dim=1)
    assert torch.equal(emb, ref)
    out = t(cat, num)
    assert out.shape == (B, 256)
    np.testing.assert_allclose(out.norm(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    state = {"user_tower." + k: v.detach().cpu().numpy() for k, v in t.state_dict().items()}
    sel = slice(0, 64)
    ref_out = otowers.tower_forward(state, "user_tower", cat[sel].cpu().numpy(), num[sel].cpu().numpy())
    assert np.abs(out[sel].cpu().numpy() - ref_out).max() < TOWER_ATOL


@pytest.mark.parametrize("cfg_name", ["cfg1", "small"])
@pytest.mark.parametrize("B", [1, 129, 1000, 20000])
def test_fused_kernel_and_layerwise_kernels_agree(built_lib, cfg_name, B):
    """The one-kernel tower (gather -> 3 chained tcgen05 GEMMs, activations in shared memory) and the
    layer-by-layer kernels compute the same 16-bit-operand / fp32-accumulate arithmetic: both within the
    tolerance of the fp32 oracle and within fp32 summation-order noise of each other."""
    import torch
    from oracle import towers as otowers
    from weights import make_inputs
    m, fx, cfg, state = _model(cfg_name)
    ucat, unum, acat = make_inputs(cfg, 99 + B, B)
    outs = {}
    for path in (1, 2):
        for tower in (m.user_tower, m.ad_tower):
            tower.force_path = path
            tower._free()
        with torch.no_grad():
            u = m.get_user_embeddings(torch.from_numpy(ucat).cuda(), torch.from_numpy(unum).cuda()).cpu().numpy()
            a = m.get_ad_embeddings(torch.from_numpy(acat).cuda()).cpu().numpy()
        lib = built_lib
        assert lib.b2r_tower_get_param(m.user_tower._handle, b"fused") == (1.0 if path == 2 else 0.0)
        outs[path] = (u, a)
    n_ref = min(B, 2000)
    ref_u = otowers.tower_forward(state, "user_tower", ucat[:n_ref], unum[:n_ref])
    ref_a = otowers.tower_forward(state, "ad_tower", acat[:n_ref])
    for path in (1, 2):
        assert np.abs(outs[path][0][:n_ref] - ref_u).max() < TOWER_ATOL
        assert np.abs(outs[path][1][:n_ref] - ref_a).max() < TOWER_ATOL
        np.testing.assert_allclose(np.linalg.norm(outs[path][0], axis=1), 1.0, atol=1e-5)
    assert np.abs(outs[1][0] - outs[2][0]).max() < 2e-6
    assert np.abs(outs[1][1] - outs[2][1]).max() < 2e-6


@pytest.mark.parametrize("path", [1, 2])
def test_fp16_saturation_is_detected_and_rerun_in_bf16(built_lib, path):
    """A checkpoint whose hidden activations leave the fp16 range (first Linear scaled x 2e5) must not be
    silently clipped: the first forward reports the saturation, the tower switches to bf16 operands
    (north_star's format) and reruns the batch; the result tracks the fp32 oracle."""
    import torch
    from movie_recommender_demo_b200.two_tower_model import TwoTowerModel
    from oracle import towers as otowers
    from weights import CONFIGS, feature_dims, make_inputs, make_state
    cfg = CONFIGS["small"]
    user, ad = feature_dims(cfg)
    state = {k: np.asarray(v).copy() for k, v in make_state(cfg, 5).items()}
    for k in ("user_tower.mlp.0.weight", "user_tower.mlp.0.bias"):
        state[k] = state[k] * np.float32(2e5)
    m = TwoTowerModel(user, ad, cfg["numerical_dim"], cfg["embedding_dim"], cfg["hidden_dims"], cfg["output_dim"])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    m = m.to("cuda").eval()
    m.user_tower.force_path = path
    ucat, unum, _ = make_inputs(cfg, 3, 300)
    ref = otowers.tower_forward(state, "user_tower", ucat, unum)
    with pytest.warns(UserWarning, match="bf16"):
        u = m.get_user_embeddings(torch.from_numpy(ucat).cuda(), torch.from_numpy(unum).cuda()).cpu().numpy()
    assert m.user_tower.native_operand_dtype == "bf16"
    assert np.abs(u - ref).max() < 2e-2          # bf16 operands: 8x the fp16 rounding error per layer
    cos = (u * ref).sum(1)
    assert cos.min() > 0.999
    # pinned to fp16 the same checkpoint clips: the flag is still raised, the answer is visibly off
    m2 = TwoTowerModel(user, ad, cfg["numerical_dim"], cfg["embedding_dim"], cfg["hidden_dims"], cfg["output_dim"])
    m2.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    m2 = m2.to("cuda").eval()
    m2.user_tower.force_path = path
    m2.user_tower.operand_dtype = "fp16"
    u16 = m2.get_user_embeddings(torch.from_numpy(ucat).cuda(), torch.from_numpy(unum).cuda()).cpu().numpy()
    assert m2.user_tower.native_operand_dtype == "fp16"
    assert np.abs(u16 - ref).max() > np.abs(u - ref).max()


def test_index_check_is_deferred_after_the_first_forward(built_lib):
    """No host sync per forward: the first forward on fresh weights is checked synchronously, later ones report
    a bad id at the next forward or in check() (torch on CUDA reports it asynchronously too)."""
    import torch
    m, fx, cfg, _ = _model("small")
    good = torch.from_numpy(fx["ucat"]).cuda()
    num = torch.from_numpy(fx["unum"]).cuda()
    ref = m.get_user_embeddings(good, num)
    bad = fx["ucat"].copy()
    bad[1, 0] = -1
    out = m.get_user_embeddings(torch.from_numpy(bad).cuda(), num)      # returns: the check is deferred
    assert out.shape == ref.shape
    with pytest.raises(IndexError):
        m.user_tower.check()
    assert torch.equal(m.get_user_embeddings(good, num), ref)            # flag was cleared
    m.get_user_embeddings(torch.from_numpy(bad).cuda(), num)
    torch.cuda.synchronize()
    with pytest.raises(IndexError, match="earlier forward"):
        m.get_user_embeddings(good, num)
    m.user_tower.sync_checks = True
    with pytest.raises(IndexError):
        m.get_user_embeddings(torch.from_numpy(bad).cuda(), num)
    emb = m.user_tower.embedding_layer
    emb(good)
    emb(torch.from_numpy(bad).cuda())
    with pytest.raises(IndexError):
        emb.check()


@pytest.mark.parametrize("B", [1, 128, 129, 255, 257, 1000, 20000, 65536])
def test_fused_tower_on_cta_pairs_matches_the_one_cta_kernel(built_lib, B):
    """`pair = 1`: two CTAs of a cluster drive one tcgen05.mma.cta_group::2 stream (256-column weight tiles, each
    CTA loads half of every tile; an odd tile count leaves rank 1 an all-padding tile).  Same arithmetic as the
    one-CTA fused kernel: equal within fp32 summation-order noise, and within tolerance of the fp32 oracle."""
    import torch
    from oracle import towers as otowers
    from weights import make_inputs
    m, fx, cfg, state = _model("cfg1")
    ucat, unum, acat = make_inputs(cfg, 7 + B, B)
    outs = {}
    for pair in (0, 1):
        for tower in (m.user_tower, m.ad_tower):
            tower.force_path = 2
            tower.pair = pair
            tower._free()
        with torch.no_grad():
            u = m.get_user_embeddings(torch.from_numpy(ucat).cuda(), torch.from_numpy(unum).cuda()).cpu().numpy()
            a = m.get_ad_embeddings(torch.from_numpy(acat).cuda()).cpu().numpy()
        assert built_lib.b2r_tower_get_param(m.user_tower._handle, b"pair") == float(pair)
        outs[pair] = (u, a)
    n_ref = min(B, 1000)
    ref_u = otowers.tower_forward(state, "user_tower", ucat[:n_ref], unum[:n_ref])
    ref_a = otowers.tower_forward(state, "ad_tower", acat[:n_ref])
    assert np.abs(outs[1][0][:n_ref] - ref_u).max() < TOWER_ATOL
    assert np.abs(outs[1][1][:n_ref] - ref_a).max() < TOWER_ATOL
    assert np.abs(outs[0][0] - outs[1][0]).max() < 2e-6
    assert np.abs(outs[0][1] - outs[1][1]).max() < 2e-6
