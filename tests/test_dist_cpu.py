"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (twopaco_b200/dist.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from twopaco_b200 import dist as tdist


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. variable-length all-gather keeps rank order
        local = torch.arange(rank * 100, rank * 100 + (3 if rank == 0 else 5), dtype=torch.int64)
        allv = tdist.allgather_varlen(local)
        assert allv.tolist() == [0, 1, 2, 100, 101, 102, 103, 104]
        assert tdist.allgather_varlen(torch.empty(0, dtype=torch.int64)).numel() == 0
        # 2. OR of disjoint masks via sum (bit 31 included: int32 wrap-around must not matter)
        m = torch.zeros(4, dtype=torch.int32)
        if rank == 0:
            m[0], m[1] = 0b0101, -2**31
        else:
            m[0], m[2] = 0b1010, 7
        tdist.or_reduce_disjoint_(m)
        assert m.tolist() == [0b1111, -2**31, 7, 0]
        m2 = torch.zeros(4, dtype=torch.int32)
        m2[rank * 2] = 5 + rank
        m2[3 - rank * 2] = 64
        tdist.or_reduce_scatter_disjoint_(m2, rank, world)
        assert m2[rank * 2:rank * 2 + 2].tolist() == ([5, 64] if rank == 0 else [6, 64])
        # 3. exclusive prefix of (records, stubs)
        before, total = tdist.exclusive_prefix([10 + rank, 1], "cpu")
        assert total == [21, 2] and before == ([0, 0] if rank == 0 else [10, 1])
        # 4. host-sliced upload + all-gather reassembles the packed genome on every rank (ragged tail)
        import numpy as np
        rng = np.random.default_rng(5)
        codes = rng.integers(0, 2**63, size=1001, dtype=np.uint64)
        nmask = rng.integers(0, 2**63, size=503, dtype=np.uint64)
        sh = tdist.host_shard(codes, nmask, 1001 * 32 - 7, None, None, rank, world, pin=False)
        assert sh.codes.numel() == (501 if rank == 0 else 500) and sh.n_mask.numel() == (252 if rank == 0 else 251)
        c, m = tdist.upload_allgather(sh, rank, world, "cpu")
        assert np.array_equal(c.numpy().view(np.uint64), codes) and np.array_equal(m.numpy().view(np.uint64), nmask)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_dist_helpers_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_bounds_cover_everything():
    for n, world in ((0, 3), (1, 4), (1001, 2), (8 * 1024 + 8, 8), (7, 8)):
        b = [tdist.shard_bounds(n, r, world) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(hi - lo <= tdist.shard_chunk(n, world) for lo, hi in b)


@pytest.mark.parametrize("npos,world", [(1, 1), (8192, 2), (100_000, 3), (21_700_000_000, 8), (5, 4)])
def test_position_cuts(npos, world):
    cuts = tdist.position_cuts(npos, world)
    assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == npos
    assert all(a <= b for a, b in zip(cuts, cuts[1:]))
    assert all(c % tdist.TILE_POSITIONS == 0 or c == npos for c in cuts[:-1])
    tiles = (npos + tdist.TILE_POSITIONS - 1) // tdist.TILE_POSITIONS
    chunk = (tiles + world - 1) // world
    assert all(b - a <= chunk * tdist.TILE_POSITIONS for a, b in zip(cuts, cuts[1:]))
