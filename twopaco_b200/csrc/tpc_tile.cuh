// tpc_tile.cuh -- CTA-tile helpers shared by the kernels that work on a SPARSE set of positions of a
// tile (owned positions: k_bin_list; candidate marks: k_insert, k_emit_*).  One thread per 32
// positions would leave most lanes idle there, so the tile's set bits are first compacted into a
// CTA-wide list in shared memory and the per-position work then runs densely, one list entry per
// thread; the tile's slice of the packed genome is staged in shared memory for it.
#pragma once
#include "tpc_device.cuh"

namespace tpc {

// ---- ownership planes: the local round (1 + round, 0 = none) that owns each position, as bit planes
constexpr int kMaxOwnPlanes = 4;   // rounds per GPU <= 15 share one ownership scan
struct OwnPlanes {
    uint32_t* p[kMaxOwnPlanes];
    uint32_t n;        // planes in use (0 = no planes available)
    uint32_t id;       // readers: the local round wanted, + 1
};

// ownership word (1 bit per position) of the wanted round out of the planes
__device__ __forceinline__ uint32_t own_word(const OwnPlanes& op, uint64_t w) {
    uint32_t own = ~0u;
#pragma unroll
    for (int j = 0; j < kMaxOwnPlanes; ++j)
        if (j < (int)op.n) {
            const uint32_t v = __ldcs(op.p[j] + w);
            own &= ((op.id >> j) & 1u) ? v : ~v;
        }
    return own;
}

// staged words of a tile: the tile + one word before + the read-ahead of the longest k-mer of W words that starts in
// the tile (W <= 4: k <= 127; W <= 19: k <= 603 -- the kernels of the short k-mers keep their shared-memory footprint)
template <int W> struct TileWords { static constexpr int code = kTileThreads + (W <= 4 ? 8 : 24), mask = kTileThreads / 2 + (W <= 4 ? 4 : 12); };
// read-ahead padding behind the last tile of the packed genome / the n-mask (tpc_code_words, tpc_mask_words): covers
// the staging above and the W + 1 words Window::load reads from the last word of the last tile
constexpr int kCodePadWords = 32;
constexpr int kMaskPadWords = 16;

template <int W>
struct TileStage {
    uint64_t codes[2][TileWords<W>::code];   // double-buffered: code words cw_base .. of the tile
    uint64_t nmask[2][TileWords<W>::mask];   // n-mask words mw_base ..
    uint16_t list[kTilePos];             // tile-local positions of the set bits, in position order
    uint32_t warp_tot[kTileThreads / 32];
    uint32_t total;
};

struct TileGeom {
    uint32_t c_off;   // tile-local position -> position inside TileStage::codes[buf]
    uint32_t m_off;   // tile-local position -> position inside TileStage::nmask[buf]
};

// 8-byte asynchronous global->shared copy (LDGSTS)
__device__ __forceinline__ void tile_cp_async8(void* smem, const void* gmem) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}

// Request the packed-genome words of `tile` into buffer `buf` (all threads of the CTA call this; one
// cp.async group per call).  tile_compact of that tile waits for them.
template <int W>
__device__ __forceinline__ void tile_request(TileStage<W>& ts, const GenomeView& g, uint64_t tile, int buf) {
    const uint64_t tw0 = tile * kTileThreads, cw_base = tw0 ? tw0 - 1 : 0, mw_base = cw_base >> 1;
    for (int j = threadIdx.x; j < TileWords<W>::code; j += kTileThreads) tile_cp_async8(&ts.codes[buf][j], g.codes + cw_base + j);
    for (int j = threadIdx.x; j < TileWords<W>::mask; j += kTileThreads) tile_cp_async8(&ts.nmask[buf][j], g.nmask + mw_base + j);
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// All threads of the CTA call this with their 32-position bit word of the tile, whose genome words were
// requested into buffer `buf` before.  `prefetch_next()` is called by every thread once the other buffer
// is free (the caller loads the next tile's bit word and calls tile_request there), so that the next
// tile's HBM latency overlaps this tile's work.  Returns the number of set bits of the tile; ts.list
// holds them; ts.codes[buf] / ts.nmask[buf] hold the tile's genome.
template <int W, typename Prefetch>
__device__ __forceinline__ uint32_t tile_compact(TileStage<W>& ts, uint64_t tile, int buf, uint32_t bits, TileGeom& tg,
                                                 Prefetch prefetch_next) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t tw0 = tile * kTileThreads;              // first code word of the tile
    const uint64_t cw_base = tw0 ? tw0 - 1 : 0;
    const uint64_t mw_base = cw_base >> 1;
    tg.c_off = (uint32_t)(tw0 - cw_base) * 32;
    tg.m_off = (uint32_t)(tw0 * 32 - mw_base * 64);
    uint32_t cnt = __popc(bits), incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // my copies of this tile's genome words have landed
    __syncthreads();  // ... everybody's; the previous tile's readers of the stage are done
    if (lane == 31) ts.warp_tot[wid] = incl;
    prefetch_next();
    __syncthreads();
    uint32_t off = incl - cnt;
    for (int j = 0; j < wid; ++j) off += ts.warp_tot[j];
    if (tid == kTileThreads - 1) ts.total = off + cnt;
    while (bits) {
        int i = __ffs(bits) - 1;
        bits &= bits - 1;
        ts.list[off++] = (uint16_t)(tid * 32 + i);
    }
    __syncthreads();
    return ts.total;
}

// k-mer at position lp of a word array in shared memory
template <int W>
__device__ __forceinline__ Kmer<W> extract_kmer_smem(const uint64_t* words, uint32_t lp, uint32_t k) {
    Kmer<W> x;
    const uint32_t wi = lp >> 5, sh = 2 * (lp & 31);
    uint64_t lo = words[wi];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        uint64_t hi = words[wi + j + 1];
        x.w[j] = (lo >> sh) | ((hi << 1) << (63 - sh));
        lo = hi;
    }
    x.w[W - 1] &= top_mask<W>(k);
    return x;
}
__device__ __forceinline__ uint32_t stage_base(const uint64_t* codes, uint32_t lp) {   // lp: position inside the staged codes
    return (uint32_t)(codes[lp >> 5] >> (2 * (lp & 31))) & 3u;
}
__device__ __forceinline__ uint32_t stage_n(const uint64_t* nmask, uint32_t mp) {      // mp: position inside the staged n-mask
    return (uint32_t)(nmask[mp >> 6] >> (mp & 63)) & 1u;
}

}  // namespace tpc
