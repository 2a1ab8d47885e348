"""Multi-GPU plumbing for the hash-range sharded run (one process per GPU, torch.distributed).

The exchanges of DESIGN.md section 4 ("Multi-GPU"), written against torch.distributed so they run
over NCCL on GPUs and over gloo on CPU (tests/test_dist_cpu.py):

  1. all-gather of the shards' junction words (variable length)      -> identical global id index
  2. OR-reduce of the disjoint candidate masks (sum == or)
  3. exclusive prefix of (records, stubs) over the position slices    -> ordered, position-sharded emit
"""
from __future__ import annotations

import torch
import torch.distributed as dist

TILE_POSITIONS = 8192


def position_cuts(n_positions: int, world: int) -> list[int]:
    """Slice boundaries (multiples of the 8192-position tile, last = n_positions): rank r emits
    positions [cuts[r], cuts[r+1])."""
    tiles = (n_positions + TILE_POSITIONS - 1) // TILE_POSITIONS
    return [min(n_positions, (tiles * r // world) * TILE_POSITIONS) for r in range(world)] + [n_positions]


def allgather_varlen(local: torch.Tensor) -> torch.Tensor:
    """Concatenation over ranks (in rank order) of 1-D tensors of different lengths."""
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    counts[rank] = local.numel()
    dist.all_reduce(counts)
    counts_h = counts.tolist()
    mx = max(max(counts_h), 1)
    padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    gathered = [torch.empty(mx, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(gathered, padded)
    return torch.cat([gathered[r][:counts_h[r]] for r in range(world)])


def or_reduce_disjoint_(mask_words: torch.Tensor) -> torch.Tensor:
    """In-place OR over ranks of bit masks whose set bits are disjoint between ranks (each position's
    k-mer belongs to exactly one hash range, vertexenumerator.h:638), so integer sum == OR."""
    dist.all_reduce(mask_words, op=dist.ReduceOp.SUM)
    return mask_words


def exclusive_prefix(values: list[int], device) -> tuple[list[int], list[int]]:
    """Each rank contributes a small vector; returns (sum over lower ranks, sum over all ranks)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    t = torch.zeros(world, len(values), dtype=torch.int64, device=device)
    t[rank] = torch.tensor(values, dtype=torch.int64, device=device)
    dist.all_reduce(t)
    h = t.tolist()
    before = [sum(h[r][j] for r in range(rank)) for j in range(len(values))]
    total = [sum(h[r][j] for r in range(world)) for j in range(len(values))]
    return before, total


def sharded_run(session, genome, rank: int, world: int, out=None):
    """One full pass of the path on this rank's hash-range shard (`session` was created with
    shard_index=rank, shard_count=world and already holds the genome).  Returns
    (info dict, device output buffer holding this rank's contiguous slice of the de_bruijn.bin
    image at bytes [slice_offset, slice_offset + slice_bytes))."""
    from . import api
    s = session
    s.find_candidates()
    ptr, n = s.local_junctions()
    if world == 1:
        s.set_junctions(ptr, n)
        cut = [0, genome.n_positions]
        nrec, nstub = s.emit_count(0, genome.n_positions)
        rb = sb = 0
        trec, tstub, nj = nrec, nstub, n
    else:
        allj = allgather_varlen(api.as_torch(ptr, n, torch.int64))            # exchange 1
        s.set_junctions(allj.data_ptr(), allj.numel())
        mptr, mw = s.candidate_mask()
        or_reduce_disjoint_(api.as_torch(mptr, mw, torch.int32))               # exchange 2
        cut = position_cuts(genome.n_positions, world)
        nrec, nstub = s.emit_count(cut[rank], cut[rank + 1])
        (rb, sb), (trec, tstub) = exclusive_prefix([nrec, nstub], "cuda")     # exchange 3
        nj = allj.numel()
    need = 12 * (nrec + len(genome.rec_len)) + 16
    if out is None or out.nbytes < need:
        out = api.DeviceBuffer(need)
    off, nb = s.emit_write(rb, sb, out.ptr, out.nbytes)
    info = dict(junctions=nj, records=trec, stubs=tstub, slice_offset=off, slice_bytes=nb,
                image_bytes=nb if world == 1 else None)
    return info, out
