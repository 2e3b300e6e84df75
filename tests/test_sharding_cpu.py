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


# ---- the packed exchange (sharded.ShardedFlatIndex._exchange): host logic on CPU tensors under gloo ------------
class _CpuShard:
    """ShardedFlatIndex with the two CUDA kernels of the exchange (b2r_topk_pack / b2r_topk_merge_packed,
    include/b2retr.h) replaced by numpy restatements of their documented contract, so that the layouts, the
    collectives, the query slicing and the status OR run on CPU tensors under gloo."""

    @staticmethod
    def make(total_rows, k, largest=1):
        from movie_recommender_demo_b200.sharded import ShardedFlatIndex, shard_rows
        obj = ShardedFlatIndex.__new__(ShardedFlatIndex)
        obj._torch = torch
        obj.group = None
        obj.world, obj.rank = dist.get_world_size(), dist.get_rank()
        obj.total_rows = total_rows
        obj.lo, obj.hi = shard_rows(total_rows, obj.world, obj.rank)
        obj._largest = largest
        obj._bases = None
        obj._graphs = {}

        def pack(q, q_rows, k, Dl, Il, st, out):
            o = out.numpy()
            o[:q, :k] = Dl.numpy().view(np.int32)
            lab = Il.numpy()
            o[:q, k:2 * k] = np.where(lab < 0, -1, lab - obj.lo).astype(np.int32)
            o[:q, 2 * k] = 0 if st is None else st.numpy()
            o[q:, :k] = np.float32(-3.4028234663852886e38 if largest else 3.4028234663852886e38).view(np.int32)
            o[q:, k:2 * k] = -1
            o[q:, 2 * k] = 0

        def merge(P, q, q_stride, k, packed, D_out, I_out, st_out):
            pk = packed.numpy().reshape(P, q_stride, 2 * k + 1)[:, :q]
            bases = np.array([shard_rows(total_rows, obj.world, r)[0] for r in range(P)], dtype=np.int64)
            sc = pk[:, :, :k].copy().view(np.float32)
            lab = pk[:, :, k:2 * k].astype(np.int64)
            glob = np.where(lab < 0, -1, lab + bases[:, None, None])
            for qi in range(q):
                s = sc[:, qi].reshape(-1)
                g = glob[:, qi].reshape(-1)
                order = np.lexsort((np.arange(s.size), -s if largest else s))[:k]   # (score, shard, position)
                D_out.numpy()[qi] = s[order]
                I_out.numpy()[qi] = g[order]
                st_out.numpy()[qi] = np.bitwise_or.reduce(pk[:, qi, 2 * k])

        obj._pack, obj._merge = pack, merge
        return obj


def _exchange_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from movie_recommender_demo_b200.sharded import exchange_mode, shard_rows, slice_rows
        from oracle.flat import OracleIndexFlatIP, normalize_L2
        N, d = 4001, 24
        rng = np.random.default_rng(0)
        x = normalize_L2(rng.standard_normal((N, d)).astype(np.float32))
        x[1500:1510] = x[7]                                     # exact ties across the shard boundary region
        full = OracleIndexFlatIP(d)
        full.add(x)
        lo, hi = shard_rows(N, world, rank)
        local = OracleIndexFlatIP(d)
        local.add(x[lo:hi])
        for Q, k, mode in ((40, 33, "gather"), (301, 50, "sliced"), (3, 2100, "gather")):
            assert exchange_mode(Q, world) == mode
            q = normalize_L2(rng.standard_normal((Q, d)).astype(np.float32))
            q[0] = x[7]
            Dl, Il = local.search(q, k)
            Il = np.where(Il >= 0, Il + lo, -1)
            st = np.zeros(Q, dtype=np.int32)
            st[rank::5] = 1 << rank                              # each shard flags its own pattern of queries
            sh = _CpuShard.make(N, k)
            D, I, st_all = sh._exchange(torch.from_numpy(Dl), torch.from_numpy(Il), torch.from_numpy(st), k)
            Df, If = full.search(q, k)
            assert D.shape == (Q, k) and I.shape == (Q, k) and st_all.shape == (Q,)
            assert np.array_equal(I.numpy(), If), f"{mode}: sharded ids differ from the unsharded search"
            np.testing.assert_array_equal(D.numpy(), Df)
            want = np.zeros(Q, dtype=np.int32)
            for r in range(world):
                want[r::5] |= 1 << r
            assert np.array_equal(st_all.numpy(), want), "status words must be OR-ed across shards"
            if mode == "sliced":
                assert slice_rows(Q, world) * world >= Q
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_packed_exchange_both_modes_gloo(tmp_path, world):
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_exchange_mode_and_slices():
    from movie_recommender_demo_b200.sharded import exchange_mode, slice_rows
    assert exchange_mode(64, 8) == "gather" and exchange_mode(128, 8) == "gather"
    assert exchange_mode(4096, 8) == "sliced" and exchange_mode(129, 2) == "sliced"
    assert exchange_mode(1, 2) == "gather"
    assert slice_rows(4096, 8) == 512 and slice_rows(301, 2) == 151 and slice_rows(5, 8) == 1
