"""Flat-import shim: put `<repo>/dropin` ahead of the reference checkout on sys.path and the
reference's `from transformer_ranker import TransformerRanker` resolves to the B200 implementation."""
import sys as _sys
from pathlib import Path as _Path

_root = str(_Path(__file__).resolve().parent.parent)
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from movie_recommender_demo_b200.transformer_ranker import TransformerRanker, fold_ranker_weights  # noqa: F401,E402

__all__ = ["TransformerRanker", "fold_ranker_weights"]
