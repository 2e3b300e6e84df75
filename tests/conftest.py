import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def built_lib():
    """libb2retr.so, (re)built in-tree when missing or older than a source / header (incremental: a no-op when
    up to date; nvcc cross-compiles without a GPU) - a stale library must never be what the tests load."""
    from movie_recommender_demo_b200 import _lib
    from movie_recommender_demo_b200.build import build
    build()
    return _lib.load()
