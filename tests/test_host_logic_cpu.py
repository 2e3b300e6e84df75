"""Host-side logic that needs no GPU."""
import pytest


@pytest.mark.parametrize("nq", [1, 1023, 1024, 2048, 2049, 4096, 5000, 8192, 20000, 65536])
def test_pipe_chunks_tile_the_batch_and_shrink_towards_the_tail(nq):
    from movie_recommender_demo_b200.faiss_retrieval import _PIPE_CHUNK, _pipe_chunks
    ch = _pipe_chunks(nq)
    assert ch[0][0] == 0 and ch[-1][1] == nq
    assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
    sizes = [hi - lo for lo, hi in ch]
    assert all(s > 0 for s in sizes)
    assert sizes[-1] <= _PIPE_CHUNK                       # the un-hidden tail copy stays small
    assert all(a <= 4 * b for a, b in zip(sizes, sizes[1:]))   # each copy hides under the next chunk's search


def test_dimension_is_validated_before_any_gpu_work():
    """Constructor argument checks come first, so they also fire on a CPU-only box (no silent fallback:
    a valid dimension then fails on the missing GPU)."""
    from movie_recommender_demo_b200 import _lib
    from movie_recommender_demo_b200.faiss_retrieval import FAISSIndex
    FAISSIndex.verbose = False
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises((RuntimeError, _lib.B2RError, OSError)):
        FAISSIndex(256, 'Flat')
    with pytest.raises(ValueError, match="Unknown index type"):
        FAISSIndex(256, 'Nope')


def _run_bench(*args, env_extra=None):
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, str(root / "bench.py"), *args], capture_output=True, text=True,
                          timeout=300, env=env)


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm): exactly one JSON line on
    stdout with the same metric/unit/config shape, `impl`, `cpu_baseline` and a zero-copy `e2e` block."""
    import json
    r = _run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--corpus-rows", "20000", "--batch", "64")
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("Flat IP top-500 over 20000x256") and d["config"]["k"] == 500
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_reference_arm_non_zero_ranks_exit_quietly():
    r = _run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env_extra={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    r = _run_bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and r.stdout.strip() == ""
