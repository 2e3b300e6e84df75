"""ctypes binding of libb2retr.so (C ABI in include/b2retr.h).

There is NO CPU fallback: if the library is missing or was not built for this GPU the
accessors raise.  `load()` only dlopen()s the .so (works without a GPU, used by the
symbol-export test); every compute entry point needs a B200.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libb2retr.so"
HEADER_PATH = PKG_DIR.parent / "include" / "b2retr.h"

# error codes (include/b2retr.h)
OK, EINVAL, ECUDA, ENOMEM, ESTATE, EUNSUPPORTED = 0, -1, -2, -3, -4, -5
KIND_FLAT, KIND_IVF_FLAT, KIND_IVF_PQ = 0, 1, 2
METRIC_IP, METRIC_L2 = 0, 1
ST_TOO_FEW, ST_NEED_LOWER_TAU, ST_CAND_OVERFLOW, ST_RESCORE_OVERFLOW = 1, 2, 4, 8
TOWER_BAD_INDEX, TOWER_SATURATED = 1, 2

_lib = None


class B2RError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb2retr error {code}: {msg}")
        self.code = code


class TowerWeights(C.Structure):
    _fields_ = [
        ("num_fields", C.c_int), ("emb_dim", C.c_int), ("num_numerical", C.c_int),
        ("hidden1", C.c_int), ("hidden2", C.c_int), ("out_dim", C.c_int),
        ("cards", C.POINTER(C.c_int64)),
        ("tables", C.POINTER(C.c_void_p)),
        ("w1", C.c_void_p), ("b1", C.c_void_p),
        ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p),
    ]


class TowerLayers(C.Structure):
    """struct b2r_tower_layers (include/b2retr.h): a tower with any number of hidden layers."""
    _fields_ = [
        ("num_fields", C.c_int), ("emb_dim", C.c_int), ("num_numerical", C.c_int),
        ("num_layers", C.c_int),
        ("cards", C.POINTER(C.c_int64)),
        ("tables", C.POINTER(C.c_void_p)),
        ("widths", C.POINTER(C.c_int)),
        ("w", C.POINTER(C.c_void_p)),
        ("b", C.POINTER(C.c_void_p)),
    ]


class RankerWeights(C.Structure):
    """struct b2r_ranker_weights (include/b2retr.h)."""
    _fields_ = [
        ("n_user", C.c_int), ("n_ad", C.c_int), ("emb_dim", C.c_int), ("num_numerical", C.c_int),
        ("d_model", C.c_int), ("d_ff", C.c_int), ("n_layers", C.c_int), ("n_cross", C.c_int),
        ("n_tasks", C.c_int), ("head1", C.c_int), ("head2", C.c_int),
        ("cards", C.c_void_p), ("tables", C.c_void_p),
        ("w_proj", C.c_void_p), ("b_proj", C.c_void_p),
        ("w_attn", C.c_void_p), ("b_attn", C.c_void_p),
        ("ln1_g", C.c_void_p), ("ln1_b", C.c_void_p),
        ("w_fc1", C.c_void_p), ("b_fc1", C.c_void_p),
        ("w_fc2", C.c_void_p), ("b_fc2", C.c_void_p),
        ("ln2_g", C.c_void_p), ("ln2_b", C.c_void_p),
        ("w_cross", C.c_void_p), ("b_cross", C.c_void_p),
        ("w_h1", C.c_void_p), ("b_h1", C.c_void_p),
        ("w_h2", C.c_void_p), ("b_h2", C.c_void_p),
        ("w_h3", C.c_void_p), ("b_h3", C.c_void_p),
    ]


def declared_symbols() -> list[str]:
    """Every function name include/b2retr.h declares."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen libb2retr.so and set argument/return types. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    lib_path = Path(os.environ["B2R_LIB_PATH"]) if os.environ.get("B2R_LIB_PATH") else LIB_PATH   # A/B builds (measurement only)
    if not lib_path.exists():
        raise RuntimeError(
            f"{lib_path} is missing: build it with `python -m movie_recommender_demo_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(str(lib_path))
    vp, i64, i32, sz, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_double
    sig = {
        "b2r_version": (i32, []),
        "b2r_last_error": (C.c_char_p, []),
        "b2r_index_create": (i32, [C.POINTER(vp), i32, i32, i32, i32, i32, i32, i32]),
        "b2r_index_destroy": (i32, [vp]),
        "b2r_index_reset": (i32, [vp]),
        "b2r_index_reserve": (i32, [vp, i64, vp]),
        "b2r_index_ntotal": (i64, [vp]),
        "b2r_index_is_trained": (i32, [vp]),
        "b2r_index_train": (i32, [vp, i64, vp, C.c_uint64, vp]),
        "b2r_index_add": (i32, [vp, i64, vp, i32, vp]),
        "b2r_index_set_ids": (i32, [vp, i64, vp, vp]),
        "b2r_index_set_label_base": (i32, [vp, i64]),
        "b2r_index_set_param": (i32, [vp, C.c_char_p, dbl]),
        "b2r_index_get_param": (dbl, [vp, C.c_char_p]),
        "b2r_index_search_workspace": (sz, [vp, i32, i32, i32]),
        "b2r_index_search": (i32, [vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
        "b2r_index_export_centroids": (i32, [vp, vp]),
        "b2r_index_import_centroids": (i32, [vp, vp]),
        "b2r_index_export_codebooks": (i32, [vp, vp]),
        "b2r_index_import_codebooks": (i32, [vp, vp]),
        "b2r_index_list_sizes": (i32, [vp, vp]),
        "b2r_index_get_vectors": (i32, [vp, i64, i64, vp, vp]),
        "b2r_index_get_labels": (i32, [vp, i64, i64, vp, vp]),
        "b2r_index_get_codes": (i32, [vp, i64, i64, vp, vp]),
        "b2r_index_add_codes": (i32, [vp, i64, vp, vp, vp]),
        "b2r_index_save": (i32, [vp, C.c_char_p, vp]),
        "b2r_index_load": (i32, [C.POINTER(vp), C.c_char_p, i32, vp]),
        "b2r_topk_merge": (i32, [i32, i32, i32, vp, vp, vp, vp, i32, vp]),
        "b2r_topk_pack": (i32, [i32, i32, i32, vp, vp, vp, i64, vp, i32, vp]),
        "b2r_topk_merge_packed": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]),
        "b2r_peer_create": (i32, [C.POINTER(vp), i32, i32, sz, i32]),
        "b2r_peer_destroy": (i32, [vp]),
        "b2r_peer_handle": (i32, [vp, vp]),
        "b2r_peer_connect": (i32, [vp, vp]),
        "b2r_peer_allgather": (i32, [vp, vp, sz, C.POINTER(vp), vp]),
        "b2r_peer_ack": (i32, [vp, vp]),
        "b2r_gather_concat": (i32, [vp, vp, i32, i32, vp, i64, vp, i64, vp, vp]),
        "b2r_tower_create": (i32, [C.POINTER(vp), C.POINTER(TowerWeights), i32]),
        "b2r_tower_create_layers": (i32, [C.POINTER(vp), C.POINTER(TowerLayers), i32]),
        "b2r_tower_destroy": (i32, [vp]),
        "b2r_tower_workspace": (sz, [vp, i64]),
        "b2r_tower_set_param": (i32, [vp, C.c_char_p, dbl]),
        "b2r_tower_get_param": (dbl, [vp, C.c_char_p]),
        "b2r_tower_forward": (i32, [vp, vp, vp, i64, vp, vp, vp, sz, vp]),
        "b2r_ranker_create": (i32, [C.POINTER(vp), C.POINTER(RankerWeights), i32]),
        "b2r_ranker_destroy": (i32, [vp]),
        "b2r_ranker_workspace": (sz, [vp, i64]),
        "b2r_ranker_set_param": (i32, [vp, C.c_char_p, dbl]),
        "b2r_ranker_get_param": (dbl, [vp, C.c_char_p]),
        "b2r_ranker_forward": (i32, [vp, vp, vp, vp, i64, vp, vp, vp, sz, vp]),
        "b2r_debug_scores_tc": (i32, [vp, i32, vp, i32, vp, vp, sz, vp]),
        "b2r_debug_scores_simt": (i32, [vp, i32, vp, i32, vp, vp]),
        "b2r_debug_launch_count": (i64, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b2r_last_error()
        raise B2RError(rc, msg.decode() if msg else "")


def require_cuda():
    """Import torch and insist on a CUDA device (product paths call this first)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("movie_recommender_demo_b200 needs a CUDA (B200, sm_100a) device; "
                           "there is no CPU fallback")
    return torch
