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
        # 4. chunk-interleaved host shards + chunked all-gather reassemble the packed genome on every rank
        import numpy as np
        rng = np.random.default_rng(5)
        npos = 5 * 8192 + 100                                  # 6 tiles -> 3 chunks of 2 tiles (world 2)
        codes = rng.integers(0, 2**63, size=6 * 256 + 8, dtype=np.uint64)
        nmask = rng.integers(0, 2**63, size=6 * 128 + 8, dtype=np.uint64)
        sh = tdist.host_shard(codes, nmask, npos, None, None, rank, world, pin=False, n_chunks=3)
        assert sh.plan.tile_begin == [0, 2, 4] and sh.codes.numel() == 256 + 256 + 260
        c, m, ev = tdist.upload_allgather(sh, rank, world, "cpu")
        assert ev == [] and np.array_equal(c.numpy().view(np.uint64)[:len(codes)], codes)
        assert np.array_equal(m.numpy().view(np.uint64)[:len(nmask)], nmask)
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


@pytest.mark.parametrize("npos,world,chunks", [(1, 1, 8), (8192 * 7 + 5, 2, 3), (100_000, 3, 8), (21_700_000_000, 8, 16), (5, 4, 2)])
def test_chunk_plan_covers_everything(npos, world, chunks):
    tiles = (npos + 8191) // 8192
    cw, mw = tiles * 256 + 8, tiles * 128 + 8
    plan = tdist.ChunkPlan(npos, cw, mw, world, chunks)
    assert plan.tile_begin[0] == 0 and plan.n_chunks <= max(chunks, 1)
    assert all(t % world == 0 for t in plan.tile_begin)
    for a, total in enumerate((cw, mw)):
        covered = 0
        for c in range(plan.n_chunks):
            for r in range(world):
                lo, hi = plan.part_bounds(a, c, r)
                assert lo == min(total, covered) and hi - lo <= plan.arrays[a][1][c][1]
                covered = max(covered, hi) if hi > lo else covered
            start, part = plan.arrays[a][1][c]
            covered = min(total, start + part * world)
        assert covered == total and plan.device_words(a) >= total


@pytest.mark.parametrize("npos,world", [(1, 1), (8192, 2), (100_000, 3), (21_700_000_000, 8), (5, 4)])
def test_position_cuts(npos, world):
    cuts = tdist.position_cuts(npos, world)
    assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == npos
    assert all(a <= b for a, b in zip(cuts, cuts[1:]))
    assert all(c % tdist.TILE_POSITIONS == 0 or c == npos for c in cuts[:-1])
    tiles = (npos + tdist.TILE_POSITIONS - 1) // tdist.TILE_POSITIONS
    chunk = (tiles + world - 1) // world
    assert all(b - a <= chunk * tdist.TILE_POSITIONS for a, b in zip(cuts, cuts[1:]))
