"""Flat-import shim for `from training_pipeline import build_faiss_index` (the trainers stay the reference's)."""
import sys as _sys
from pathlib import Path as _Path

_root = str(_Path(__file__).resolve().parent.parent)
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from movie_recommender_demo_b200.training_pipeline import build_faiss_index  # noqa: F401,E402
