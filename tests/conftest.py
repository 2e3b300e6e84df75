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
    """libb2retr.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from movie_recommender_demo_b200 import _lib
    if not _lib.LIB_PATH.exists():
        from movie_recommender_demo_b200.build import build
        build()
    return _lib.load()
