"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes exercise the shard
partition, the all-gather plumbing (gather_topk) and — with the oracle's merge standing in for
the CUDA merge kernel — that shard -> local top-k -> all-gather -> merge equals the unsharded
oracle result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_rows_tile_the_corpus():
    from movie_recommender_demo_b200.sharded import shard_rows
    for total in (0, 1, 7, 1000, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, d, Q, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from movie_recommender_demo_b200.sharded import gather_topk, shard_rows
        from oracle.flat import OracleIndexFlatIP, normalize_L2, topk_desc
        rng = np.random.default_rng(0)
        x = normalize_L2(rng.standard_normal((N, d)).astype(np.float32))
        q = normalize_L2(rng.standard_normal((Q, d)).astype(np.float32))
        lo, hi = shard_rows(N, world, rank)
        local = OracleIndexFlatIP(d)
        local.add(x[lo:hi])
        Dl, Il = local.search(q, k)
        Il = np.where(Il >= 0, Il + lo, -1)                       # global labels = base + local row
        D_all, I_all = gather_topk(torch.from_numpy(Dl), torch.from_numpy(Il))
        assert D_all.shape == (world, Q, k) and I_all.shape == (world, Q, k)
        assert torch.equal(D_all[rank], torch.from_numpy(Dl))     # rank-major layout
        # merge (oracle stand-in for b2r_topk_merge): best-first over the P*k pooled entries
        pooled_D = D_all.permute(1, 0, 2).reshape(Q, world * k).numpy()
        pooled_I = I_all.permute(1, 0, 2).reshape(Q, world * k).numpy()
        Dm, pos = topk_desc(np.where(pooled_I >= 0, pooled_D, -np.inf).astype(np.float32), k)
        Im = np.take_along_axis(pooled_I, pos, axis=1)
        full = OracleIndexFlatIP(d)
        full.add(x)
        Df, If = full.search(q, k)
        assert np.array_equal(Im, If), "sharded merge differs from the unsharded search"
        np.testing.assert_array_equal(Dm, Df)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_merge(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 3001, 32, 5, 40, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
