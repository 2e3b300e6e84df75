"""The flat-import shims of INTEGRATION.md §2: with `<repo>/dropin` first on sys.path the reference's own import
lines (`from faiss_retrieval import FAISSIndex, TwoStageRetriever`, `from two_tower_model import TwoTowerModel`,
`from transformer_ranker import TransformerRanker`, inference.py:15-18) resolve to the B200 modules."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_flat_imports_resolve_to_the_b200_modules():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from faiss_retrieval import FAISSIndex, TwoStageRetriever\n"
        "from two_tower_model import TwoTowerModel, UserTower, AdTower, EmbeddingLayer\n"
        "from transformer_ranker import TransformerRanker\n"
        "from inference import AdRecommenderInference\n"
        "mods = {c.__module__ for c in (FAISSIndex, TwoStageRetriever, TwoTowerModel, TransformerRanker, AdRecommenderInference)}\n"
        "assert all(m.startswith('movie_recommender_demo_b200.') for m in mods), mods\n"
        "print('ok')\n" % str(ROOT / "dropin"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
