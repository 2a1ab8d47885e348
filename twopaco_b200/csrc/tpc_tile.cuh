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

constexpr int kTileCodeWords = kTileThreads + 8;        // tile + one word before + read-ahead (k <= 127)
constexpr int kTileMaskWords = kTileThreads / 2 + 4;

struct TileStage {
    uint64_t codes[kTileCodeWords];   // code words cw_base .. of the tile
    uint64_t nmask[kTileMaskWords];   // n-mask words mw_base ..
    uint16_t list[kTilePos];          // tile-local positions of the set bits, in position order
    uint32_t warp_tot[kTileThreads / 32];
    uint32_t total;
};

struct TileGeom {
    uint32_t c_off;   // tile-local position -> position inside TileStage::codes
    uint32_t m_off;   // tile-local position -> position inside TileStage::nmask
};

// All threads of the CTA call this with their 32-position bit word of the tile.  Returns the number of
// set bits of the tile; ts.list holds them; ts.codes / ts.nmask hold the tile's genome.
__device__ __forceinline__ uint32_t tile_compact(TileStage& ts, const GenomeView& g, uint64_t tile, uint32_t bits, TileGeom& tg) {
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
    __syncthreads();  // the previous tile's readers of the stage are done
    if (lane == 31) ts.warp_tot[wid] = incl;
    for (int j = tid; j < kTileCodeWords; j += kTileThreads) ts.codes[j] = __ldg(g.codes + cw_base + j);
    for (int j = tid; j < kTileMaskWords; j += kTileThreads) ts.nmask[j] = __ldg(g.nmask + mw_base + j);
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
__device__ __forceinline__ uint32_t stage_base(const TileStage& ts, uint32_t lp) {   // lp: position inside ts.codes
    return (uint32_t)(ts.codes[lp >> 5] >> (2 * (lp & 31))) & 3u;
}
__device__ __forceinline__ uint32_t stage_n(const TileStage& ts, uint32_t mp) {      // mp: position inside ts.nmask
    return (uint32_t)(ts.nmask[mp >> 6] >> (mp & 63)) & 1u;
}

}  // namespace tpc
