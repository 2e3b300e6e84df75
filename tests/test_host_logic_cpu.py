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


def test_one_randint_draw_consumes_the_stream_of_the_per_user_draws():
    """inference._stage2_batch draws the reference's random ad features (inference.py:246-248) for U users in ONE
    call: the global CPU generator must hand out the numbers the reference's per-user loop would have seen."""
    import torch
    torch.manual_seed(1234)
    per_user = torch.cat([torch.randint(0, 200, (500, 20)).long() for _ in range(7)])
    after_a = torch.randint(0, 200, (3,))
    torch.manual_seed(1234)
    one_draw = torch.randint(0, 200, (7 * 500, 20))
    after_b = torch.randint(0, 200, (3,))
    assert one_draw.dtype == torch.int64
    assert torch.equal(per_user, one_draw) and torch.equal(after_a, after_b)


def test_bench_extras_survive_a_dying_child(monkeypatch):
    """bench.run_extras_isolated: a sub-result whose child process dies (device fault) is recorded as an error and
    the remaining sub-results run in a new child; the skip list grows so that nothing runs twice."""
    import argparse
    import importlib.util
    import json
    import subprocess
    import types
    root = __import__("pathlib").Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("bench_under_test", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    calls = []

    def fake_run(cmd, **kw):
        path = cmd[cmd.index("--extras-child") + 1]
        skip = [n for n in cmd[cmd.index("--extras-skip") + 1].split(",") if n]
        calls.append(skip)
        todo = [n for n in bench.EXTRA_NAMES if n not in skip]
        if len(calls) == 1:      # dies while running the third sub-result
            state = {"done": {n: {"value": 1.0} for n in todo[:2]}, "running": todo[2]}
            open(path, "w").write(json.dumps(state))
            return types.SimpleNamespace(returncode=-6, stderr="CUDA error: unspecified launch failure")
        state = {"done": {n: {"value": 2.0} for n in todo}, "running": None}
        open(path, "w").write(json.dumps(state))
        return types.SimpleNamespace(returncode=0, stderr="")

    monkeypatch.setattr(subprocess, "run", fake_run)
    args = argparse.Namespace(no_anchor=True, batch=4096, scan_dtype="auto")
    msgs = []
    extra = bench.run_extras_isolated(args, msgs.append)
    names = [n for n in bench.EXTRA_NAMES if n != "flat_100M_one_gpu"]
    assert list(extra) == names                                   # every sub-result reported, anchor skipped
    assert extra[names[0]] == {"value": 1.0} and extra[names[1]] == {"value": 1.0}
    assert "died" in extra[names[2]]["error"] and "launch failure" in extra[names[2]]["error"]
    assert all(extra[n] == {"value": 2.0} for n in names[3:])
    assert len(calls) == 2 and set(calls[1]) == {"flat_100M_one_gpu", *names[:3]}
    assert len(msgs) == 1


def test_preprocess_user_batch_matches_the_reference_method():
    """inference.preprocess_user_batch (vectorised, dict look-ups) vs tests/golden/preprocess_users.npz = the
    reference's own preprocess_user_features (inference.py:159-197, run by make_preprocess_golden.py over fitted
    sklearn encoders / scaler): unseen values and absent keys fall back to 'missing' / 0, |x| is logged."""
    import json
    import numpy as np
    from movie_recommender_demo_b200.inference import AdRecommenderInference
    fx = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "preprocess_users.npz")
    users = json.loads(str(fx["users"]))
    classes = json.loads(str(fx["classes"]))

    class Enc:
        def __init__(self, c):
            self.classes_ = np.array(c, dtype=object)

    class Scaler:
        mean_, scale_ = fx["scaler_mean"], fx["scaler_scale"]

        def transform(self, x):
            return ((np.asarray(x, dtype=np.float64) - self.mean_) / self.scale_).astype(np.float32)

    class Pre:
        label_encoders = {c: Enc(v) for c, v in classes.items()}
        numerical_cols = [f"I{i + 1}" for i in range(13)]
        scaler = Scaler()

    inst = object.__new__(AdRecommenderInference)      # host-side method only: no CUDA, no models
    inst.preprocessor, inst._encoders = Pre(), None
    cat, num = inst.preprocess_user_batch(users)
    assert cat.dtype == __import__("torch").long and np.array_equal(cat.numpy(), fx["cat"])
    np.testing.assert_allclose(num.numpy().astype(np.float64), fx["num"], rtol=0, atol=1e-6)
    one_c, one_n = inst.preprocess_user_features(users[5])
    assert np.array_equal(one_c.numpy(), fx["cat"][5:6]) and tuple(one_n.shape) == (1, 13)
