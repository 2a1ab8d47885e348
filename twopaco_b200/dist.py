"""Multi-GPU plumbing for the hash-range sharded run (one process per GPU, torch.distributed).

The exchanges of DESIGN.md section 4 ("Multi-GPU"), written against torch.distributed so they run
over NCCL on GPUs and over gloo on CPU (tests/test_dist_cpu.py):

  1. all-gather of the shards' junction words (variable length)      -> identical global id index
  2. OR-reduce-scatter of the disjoint candidate masks (sum == or) into the position slices
  3. exclusive prefix of (records, stubs) over the position slices    -> ordered, position-sharded emit
"""
from __future__ import annotations

import torch
import torch.distributed as dist

TILE_POSITIONS = 8192


def position_cuts(n_positions: int, world: int) -> list[int]:
    """Slice boundaries (multiples of the 8192-position tile, last = n_positions): rank r emits
    positions [cuts[r], cuts[r+1]).  The slices are `world` equal chunks of whole tiles (the last ones
    may be short or empty) -- the chunks of the padded candidate mask (tpc_session_candidate_mask)."""
    tiles = (n_positions + TILE_POSITIONS - 1) // TILE_POSITIONS
    chunk = (tiles + world - 1) // world
    return [min(n_positions, r * chunk * TILE_POSITIONS) for r in range(world)] + [n_positions]


def all_gather_into(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dist.all_gather_into_tensor; gloo has no all-gather of CUDA tensors (ranks sharing one GPU in the tests):
    there every rank drops its part into a zeroed buffer and the buffers are summed."""
    if dist.get_backend() == "gloo" and src.is_cuda:
        rank, n = dist.get_rank(), src.numel()
        dst.zero_()
        dst[rank * n:(rank + 1) * n].copy_(src)
        dist.all_reduce(dst, op=dist.ReduceOp.SUM)
    else:
        dist.all_gather_into_tensor(dst, src)


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
    gathered = torch.empty(mx * world, dtype=local.dtype, device=local.device)
    all_gather_into(gathered, padded)
    return torch.cat([gathered[r * mx:r * mx + counts_h[r]] for r in range(world)])


def or_reduce_disjoint_(mask_words: torch.Tensor) -> torch.Tensor:
    """In-place OR over ranks of bit masks whose set bits are disjoint between ranks (each position's
    k-mer belongs to exactly one hash range, vertexenumerator.h:638), so integer sum == OR."""
    dist.all_reduce(mask_words, op=dist.ReduceOp.SUM)
    return mask_words


def or_reduce_scatter_disjoint_(mask_words: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Like or_reduce_disjoint_, but only chunk `rank` of the (padded, world equal chunks) mask is
    completed on this rank -- all the position-sharded emit reads.  NCCL: one reduce-scatter, which
    moves half the bytes of the all-reduce; gloo (CPU tests) has no reduce-scatter: all-reduce."""
    n = mask_words.numel()
    assert n % world == 0, (n, world)
    c = n // world
    if dist.get_backend() == "nccl":
        mine = torch.empty(c, dtype=mask_words.dtype, device=mask_words.device)
        dist.reduce_scatter_tensor(mine, mask_words, op=dist.ReduceOp.SUM)
        mask_words[rank * c:(rank + 1) * c].copy_(mine)
    else:
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
    import time
    from . import api
    s = session
    t0 = time.perf_counter()
    s.find_candidates()
    t_find = time.perf_counter()
    ptr, n = s.local_junctions()
    ms_x = {}
    if world == 1:
        s.set_junctions(ptr, n)
        cut = [0, genome.n_positions]
        nrec, nstub = s.emit_count(0, genome.n_positions)
        rb = sb = 0
        trec, tstub, nj = nrec, nstub, n
    else:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if torch.cuda.is_available() else None
        if ev: ev[0].record()
        allj = allgather_varlen(api.as_torch(ptr, n, torch.int64))            # exchange 1
        if ev: ev[1].record()
        s.set_junctions(allj.data_ptr(), allj.numel())
        mptr, mw = s.candidate_mask()
        if ev: ev[2].record()
        or_reduce_scatter_disjoint_(api.as_torch(mptr, mw, torch.int32), rank, world)   # exchange 2
        if ev: ev[3].record()
        cut = position_cuts(genome.n_positions, world)
        nrec, nstub = s.emit_count(cut[rank], cut[rank + 1])
        (rb, sb), (trec, tstub) = exclusive_prefix([nrec, nstub], "cuda")     # exchange 3
        nj = allj.numel()
        if ev:   # (emit_count has synchronised the stream)
            ms_x = {"ms_allgather_junctions": round(ev[0].elapsed_time(ev[1]), 3),
                    "ms_reduce_scatter_masks": round(ev[2].elapsed_time(ev[3]), 3)}
    t_count = time.perf_counter()
    need = 12 * (nrec + len(genome.rec_len)) + 16
    if out is None or out.nbytes < need:
        out = api.DeviceBuffer(need)
    off, nb = s.emit_write(rb, sb, out.ptr, out.nbytes)
    t_end = time.perf_counter()
    # host wall clock of the phases (find_candidates and the count exchange end with a synchronisation; emit_write is enqueued)
    ms_x.update(wall_ms_find_candidates=round((t_find - t0) * 1e3, 3), wall_ms_exchange_index_count=round((t_count - t_find) * 1e3, 3),
                wall_ms_emit_write_enqueue=round((t_end - t_count) * 1e3, 3))
    info = dict(junctions=nj, records=trec, stubs=tstub, slice_offset=off, slice_bytes=nb,
                image_bytes=nb if world == 1 else None, **ms_x)
    return info, out


# ---------------------------------------------------------------------------------------------
# host buffers -> host buffers on N GPUs (the multi-GPU `e2e` region of bench.py)
# ---------------------------------------------------------------------------------------------
CODE_WORDS_PER_TILE = 256      # 8192 positions x 2 bits / 64
MASK_WORDS_PER_TILE = 128


class ChunkPlan:
    """Geometry of the chunked multi-GPU upload of a packed genome of `code_words` / `mask_words` 64-bit
    words: the tiles are cut into (at most) `n_chunks` chunks of a multiple of `world` tiles; every chunk of
    both arrays is `world` equal parts, part r held (and uploaded) by rank r.  One all-gather per chunk and
    array then lands the chunk contiguously, in place, in every GPU's copy of the genome -- chunk by chunk in
    position order, so the first pass over the genome can start on the chunks that have arrived
    (tpc_session_add_genome_event).  The last chunk also carries the arrays' read-ahead padding words and is
    padded to a multiple of `world` words (the device arrays are allocated that much longer)."""

    def __init__(self, n_positions: int, code_words: int, mask_words: int, world: int, n_chunks: int = 8):
        self.world = world
        self.tiles = (n_positions + TILE_POSITIONS - 1) // TILE_POSITIONS
        per = max(1, -(-self.tiles // max(n_chunks, 1)))
        per = -(-per // world) * world                       # tiles per chunk: a multiple of world
        self.tile_begin = list(range(0, max(self.tiles, 1), per))
        self.n_chunks = len(self.tile_begin)
        self.arrays = []                                     # per array: (total words, [(start, part_words)] per chunk)
        for total, wpt in ((code_words, CODE_WORDS_PER_TILE), (mask_words, MASK_WORDS_PER_TILE)):
            chunks = []
            for c, tb in enumerate(self.tile_begin):
                start = tb * wpt
                end = total if c + 1 == self.n_chunks else self.tile_begin[c + 1] * wpt
                chunks.append((start, -(-(end - start) // world)))
            self.arrays.append((total, chunks))

    def device_words(self, a: int) -> int:
        start, part = self.arrays[a][1][-1]
        return start + part * self.world

    def part_bounds(self, a: int, c: int, rank: int) -> tuple[int, int]:
        """Word range [lo, hi) of array a that rank `rank` holds of chunk c (clipped to the array)."""
        total, chunks = self.arrays[a]
        start, part = chunks[c]
        return min(total, start + rank * part), min(total, start + (rank + 1) * part)

    def host_words(self, a: int) -> int:
        """Length of a rank's host buffer for array a: its (padded) parts of all chunks, back to back."""
        return sum(part for _, part in self.arrays[a][1])


class HostGenomeShard:
    """What rank r holds of a packed genome in (pinned) host memory: for every chunk of the ChunkPlan its
    1/world part of the codes and of the n_mask, back to back, plus the (small) record table of the whole
    input.  Every rank uploads only this over its own PCIe link; the parts are all-gathered over NVLink
    (`upload_allgather`), which replaces the reference's per-stage re-read of the whole FASTA by every
    worker (vertexenumerator.h:1135-1214)."""

    def __init__(self, plan: ChunkPlan, codes: torch.Tensor, n_mask: torch.Tensor, n_positions: int, rec_start, rec_len):
        self.plan, self.codes, self.n_mask = plan, codes, n_mask           # int64 tensors (pinned when possible)
        self.n_positions, self.rec_start, self.rec_len = int(n_positions), rec_start, rec_len

    @property
    def nbytes(self) -> int:
        return (self.codes.numel() + self.n_mask.numel()) * 8


def host_shard(codes, n_mask, n_positions: int, rec_start, rec_len, rank: int, world: int, pin: bool = True,
               n_chunks: int = 8, fetch=None) -> HostGenomeShard:
    """Rank `rank`'s shard of numpy uint64 arrays `codes` / `n_mask` (or of any source `fetch(a, lo, hi)` ->
    numpy uint64 words [lo, hi) of array a, e.g. a device-resident genome)."""
    import numpy as np
    plan = ChunkPlan(n_positions, len(codes) if fetch is None else codes, len(n_mask) if fetch is None else n_mask, world, n_chunks)
    if fetch is None:
        src = (codes, n_mask)
        fetch = lambda a, lo, hi: src[a][lo:hi]
    out = []
    for a in range(2):
        buf = np.zeros(plan.host_words(a), dtype=np.uint64)
        off = 0
        for c in range(plan.n_chunks):
            lo, hi = plan.part_bounds(a, c, rank)
            buf[off:off + hi - lo] = fetch(a, lo, hi)
            off += plan.arrays[a][1][c][1]
        t = torch.from_numpy(buf.view(np.int64))
        out.append(t.pin_memory() if pin and torch.cuda.is_available() else t)
    return HostGenomeShard(plan, out[0], out[1], n_positions, rec_start, rec_len)


def upload_allgather(shard: HostGenomeShard, rank: int, world: int, device):
    """-> (codes, n_mask, events): the WHOLE genome as int64 tensors on `device`, produced chunk by chunk
    on a side stream: H2D of this rank's part of the chunk, then one all-gather per array (NCCL over NVLink
    on GPUs, gloo in the CPU tests) straight into place.  events[c] = (first tile of chunk c, CUDA event
    recorded when chunk c is complete) -- empty on CPU, where everything is synchronous."""
    plan = shard.plan
    on_gpu = torch.device(device).type == "cuda"
    full = [torch.empty(plan.device_words(a), dtype=torch.int64, device=device) for a in range(2)]
    # two staging buffers per array: the H2D copy of chunk c+1 (copy engine, PCIe) runs beside the all-gather of
    # chunk c (NVLink), each on its own side stream
    stage = [[torch.empty(max(p for _, p in plan.arrays[a][1]), dtype=torch.int64, device=device) for _ in range(2)] for a in range(2)]
    events = []
    import contextlib
    h2d = torch.cuda.Stream() if on_gpu else None
    ag = torch.cuda.Stream() if on_gpu else None
    on = lambda st: torch.cuda.stream(st) if on_gpu else contextlib.nullcontext()
    if on_gpu:
        h2d.wait_stream(torch.cuda.current_stream())       # the allocations above
        ag.wait_stream(torch.cuda.current_stream())
    gathered = [[None, None], [None, None]]                # all-gather that last read staging buffer [a][b]
    offs = [0, 0]
    for c in range(plan.n_chunks):
        for a, host in enumerate((shard.codes, shard.n_mask)):
            start, part = plan.arrays[a][1][c]
            b = c & 1
            with on(h2d):
                if on_gpu and gathered[a][b] is not None:
                    h2d.wait_event(gathered[a][b])
                mine = full[a][start + rank * part:start + (rank + 1) * part] if world == 1 else stage[a][b][:part]
                mine.copy_(host[offs[a]:offs[a] + part], non_blocking=True)
                offs[a] += part
                if on_gpu:
                    copied = torch.cuda.Event()
                    copied.record(h2d)
            with on(ag):
                if on_gpu:
                    ag.wait_event(copied)
                if world > 1:
                    all_gather_into(full[a][start:start + part * world], mine)
                if on_gpu:
                    gathered[a][b] = torch.cuda.Event()
                    gathered[a][b].record(ag)
        if on_gpu:
            events.append((plan.tile_begin[c], gathered[1][c & 1]))   # n_mask of chunk c is gathered last
    if on_gpu:
        for t in full + stage[0] + stage[1]:
            t.record_stream(h2d)
            t.record_stream(ag)
    return full[0], full[1], events


def sharded_run_host(shard: HostGenomeShard, rank: int, world: int, k: int, filter_bits: int, q: int = 5, rounds: int = 1,
                     out_host: torch.Tensor | None = None, dev_out=None):
    """Host buffers in, host buffers out, on `world` GPUs: chunked upload + all-gather of the packed genome
    (overlapped with the first pass over it), the sharded run, and the device->host copy of this rank's
    slice of the de_bruijn.bin image into `out_host` (uint8, pinned; bytes [0, slice_bytes) = image bytes
    [slice_offset, +slice_bytes) -- a rank would pwrite() them at that offset).
    Returns (info, out_host, dev_out)."""
    from . import api
    codes, n_mask, events = upload_allgather(shard, rank, world, "cuda")
    s = api.Session(k=k, filter_bits=filter_bits, q=q, rounds=rounds, shard_index=rank, shard_count=world)
    try:
        s.set_genome_device(codes.data_ptr(), n_mask.data_ptr(), shard.n_positions, shard.rec_start, shard.rec_len,
                            keep=(codes, n_mask))
        for tile_begin, ev in events:
            s.add_genome_event(tile_begin, ev.cuda_event, keep=ev)

        class _G:  # what sharded_run needs to know about the genome
            n_positions, rec_len = shard.n_positions, shard.rec_len
        info, dev_out = sharded_run(s, _G, rank, world, dev_out)
        nb = info["slice_bytes"]
        if out_host is None or out_host.numel() < nb:
            out_host = torch.empty(max(nb, 16), dtype=torch.uint8)
            if torch.cuda.is_available():
                out_host = out_host.pin_memory()
        if nb:
            out_host[:nb].copy_(api.as_torch(dev_out.ptr, nb, torch.uint8), non_blocking=True)
        torch.cuda.synchronize()
        info["stats"] = s.stats()
    finally:
        s.close()
    return info, out_host, dev_out
