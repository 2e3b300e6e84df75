"""Flat-import shim: put `<repo>/dropin` ahead of the reference checkout on sys.path and the
reference's `from two_tower_model import ...` lines resolve to the B200 implementation."""
import sys as _sys
from pathlib import Path as _Path

_root = str(_Path(__file__).resolve().parent.parent)
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from movie_recommender_demo_b200.two_tower_model import *  # noqa: F401,F403,E402
from movie_recommender_demo_b200.two_tower_model import __all__  # noqa: F401,E402
