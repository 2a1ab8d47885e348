// tpc_device.cuh -- device-side primitives shared by the junction-finding kernels (sm_100a).
//
// Data layout in HBM:
//   codes   : 2 bits / position, position p -> bits 2(p%32) of 64-bit word p/32 (A0 C1 G2 T3)
//             == the reference's CompressedString layout (compressedstring.h:239-264) applied
//             to the whole genome.
//   n_mask  : 1 bit / position (1 = 'N' / record separator / padding).
//   k-mers  : W = ceil(k/32) 64-bit words, base j of the k-mer at bits 2(j%32) of word j/32
//             ("LSB-first").  X = k-mer as written, Y = its reverse complement.  canonical =
//             min(X, Y) compared from the most significant word.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tpc {

constexpr int kTileThreads = 256;
constexpr int kPosPerThread = 32;
constexpr int kTilePos = kTileThreads * kPosPerThread;  // 8192 positions per CTA tile

constexpr int kPosBits = 40;
constexpr uint64_t kPosMask = (1ull << kPosBits) - 1;

// meta word of a candidate-table slot (canonical orientation of the k-mer)
constexpr uint64_t kMetaInN1 = 1ull << 8;    // an 'N' / sequence end seen on the in side
constexpr uint64_t kMetaInN2 = 1ull << 9;    // ... seen at least twice (every N is unique)
constexpr uint64_t kMetaOutN1 = 1ull << 10;
constexpr uint64_t kMetaOutN2 = 1ull << 11;
constexpr int kMetaCountShift = 16;          // occurrence count (only when -a is given)

template <int W>
struct Kmer {
    uint64_t w[W];
};

struct GenomeView {
    const uint64_t* __restrict__ codes;
    const uint64_t* __restrict__ nmask;
    uint64_t npos;
};

struct Slot {  // candidate table T and junction index J
    unsigned long long rep;   // 0 = empty, else tag(24) << 40 | position(40) of an occurrence
    unsigned long long meta;  // T: neighbour flags | count << 16      J: junction id
};

struct TableView {
    Slot* slots;
    uint32_t log2cap;
    // Inline keys (k <= 31 only): the canonical k-mer (< 2^62) fits the slot, so lookups never
    // touch the genome.  T slot = {key + 1, flags(0..11) | (first position << 1 | its k-mer is on the canonical strand) << 23}
    //                    J slot = {key + 1 | first occurrence is on the canonical strand << 63, id}
    // (the J slot's first word is also the "junction key" exchanged beside the position when the genome is windowed)
    uint32_t inline_keys;
};
constexpr int kInlinePosShift = 23;

struct KParams {
    uint32_t k;
    uint32_t q;
    uint32_t sector_shift;   // 64 - log2(#sectors)
    uint32_t nparts;         // rounds * shard_count
    uint32_t part;           // part processed by this launch
    uint32_t count_occurrences;
    uint64_t seed;
};

__device__ __forceinline__ uint64_t fmix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85ebca6bu;
    x ^= x >> 13; x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}

// reverse the order of the 32 two-bit groups of a word
__device__ __forceinline__ uint64_t pairrev64(uint64_t x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

template <int W>
__device__ __forceinline__ uint64_t top_mask(uint32_t k) {
    uint32_t bits = 2 * k - 64 * (W - 1);  // 2..62 (k odd, W = ceil(k/32))
    return (~0ull) >> (64 - bits);
}

__device__ __forceinline__ uint32_t load_base(const uint64_t* __restrict__ codes, uint64_t p) {
    return (uint32_t)(__ldg(codes + (p >> 5)) >> (2 * (p & 31))) & 3u;
}

__device__ __forceinline__ uint32_t load_n(const uint64_t* __restrict__ nmask, uint64_t p) {
    return (uint32_t)(__ldg(nmask + (p >> 6)) >> (p & 63)) & 1u;
}

// 64 n-mask bits starting at position p
__device__ __forceinline__ uint64_t load_nbits64(const uint64_t* __restrict__ nmask, uint64_t p) {
    uint64_t wi = p >> 6;
    uint32_t sh = (uint32_t)(p & 63);
    uint64_t lo = __ldg(nmask + wi);
    uint64_t hi = __ldg(nmask + wi + 1);
    return (lo >> sh) | ((hi << 1) << (63 - sh));
}

// any 'N' among positions [p, p+len) ?
__device__ __forceinline__ bool any_n(const uint64_t* __restrict__ nmask, uint64_t p, uint32_t len) {
    uint64_t acc = 0;
    for (uint32_t off = 0; off < len; off += 64) {
        uint64_t b = load_nbits64(nmask, p + off);
        uint32_t rem = len - off;
        if (rem < 64) b &= (~0ull) >> (64 - rem);
        acc |= b;
    }
    return acc != 0;
}

template <int W>
__device__ __forceinline__ Kmer<W> extract_kmer(const uint64_t* __restrict__ codes, uint64_t p, uint32_t k) {
    Kmer<W> x;
    uint64_t wi = p >> 5;
    uint32_t sh = 2 * (uint32_t)(p & 31);
    uint64_t lo = __ldg(codes + wi);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        uint64_t hi = __ldg(codes + wi + j + 1);
        x.w[j] = (lo >> sh) | ((hi << 1) << (63 - sh));
        lo = hi;
    }
    x.w[W - 1] &= top_mask<W>(k);
    return x;
}

template <int W>
__device__ __forceinline__ Kmer<W> revcomp(const Kmer<W>& x, uint32_t k) {
    uint64_t t[W];
#pragma unroll
    for (int j = 0; j < W; ++j) t[j] = pairrev64(~x.w[W - 1 - j]);
    uint32_t s = 64 * W - 2 * k;  // 2..62
    Kmer<W> y;
#pragma unroll
    for (int j = 0; j < W - 1; ++j) y.w[j] = (t[j] >> s) | (t[j + 1] << (64 - s));
    y.w[W - 1] = t[W - 1] >> s;
    return y;
}

// advance both strands by one base: X drops its first base and appends `next`
template <int W>
__device__ __forceinline__ void roll(Kmer<W>& x, Kmer<W>& y, uint32_t next, uint32_t k) {
    uint32_t topbits = 2 * k - 64 * (W - 1);
#pragma unroll
    for (int j = 0; j < W - 1; ++j) x.w[j] = (x.w[j] >> 2) | (x.w[j + 1] << 62);
    x.w[W - 1] = (x.w[W - 1] >> 2) | ((uint64_t)next << (topbits - 2));
#pragma unroll
    for (int j = W - 1; j > 0; --j) y.w[j] = (y.w[j] << 2) | (y.w[j - 1] >> 62);
    y.w[0] = (y.w[0] << 2) | (uint64_t)(3u - next);
    y.w[W - 1] &= (~0ull) >> (64 - topbits);
}

template <int W>
__device__ __forceinline__ bool kmer_less(const Kmer<W>& a, const Kmer<W>& b) {
#pragma unroll
    for (int j = W - 1; j > 0; --j)
        if (a.w[j] != b.w[j]) return a.w[j] < b.w[j];
    return a.w[0] < b.w[0];
}

// c ? a : b, word by word (a ternary on the structs would force both into local memory)
template <int W>
__device__ __forceinline__ Kmer<W> kmer_select(bool c, const Kmer<W>& a, const Kmer<W>& b) {
    Kmer<W> r;
#pragma unroll
    for (int j = 0; j < W; ++j) r.w[j] = c ? a.w[j] : b.w[j];
    return r;
}

template <int W>
__device__ __forceinline__ bool kmer_eq(const Kmer<W>& a, const Kmer<W>& b) {
    bool e = true;
#pragma unroll
    for (int j = 0; j < W; ++j) e = e && (a.w[j] == b.w[j]);
    return e;
}

template <int W>
__device__ __forceinline__ uint64_t kmer_hash(const Kmer<W>& c, uint64_t seed) {
    uint64_t h = seed ^ 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int j = 0; j < W; ++j) h = fmix64(h ^ c.w[j]);
    return h;
}

// ---- hash-range ownership (rounds x GPUs).  It is evaluated for EVERY position on every GPU (k_own), so it
// must cost little; the full 64-bit hash is only needed for the positions a GPU owns.  Any function of the
// vertex that is the same for both strands will do.  k is odd, so the middle 11 bases of the reverse
// complement are the reverse complement of the middle 11 bases: the key is the smaller of the two 22-bit
// values, multiplied by a constant -- two funnel shifts, a min and a multiply per position instead of
// building, comparing and folding both 2k-bit strands.  4^11 / 2 classes spread evenly over any sensible number
// of parts (uniform to a fraction of a per cent on the synthetic sets; all copies of a low-complexity k-mer go
// to one owner under any scheme).  k < 11: multiplicative fold of the whole canonical k-mer.
constexpr uint32_t kOwnMid = 11;
constexpr uint32_t kOwnMidMask = (1u << (2 * kOwnMid)) - 1u;
__device__ __forceinline__ uint32_t pairrev32(uint32_t x) {
    x = __brev(x);
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}
// m, y: the middle 11-mer and its reverse complement, top-aligned (<< 10) -- both in any order
__device__ __forceinline__ uint32_t owner_fold_mid(uint32_t m_top, uint32_t y_top) {
    return min(m_top, y_top) * 0x9E3779B1u;
}
template <int W>
__device__ __forceinline__ uint32_t owner_fold(const Kmer<W>& canon, uint32_t k) {
    if (k >= kOwnMid) {
        const uint32_t bit = 2 * ((k - kOwnMid) >> 1), wi = bit >> 6, sh = bit & 63;
        uint64_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (j == (int)wi) lo = canon.w[j];
            if (j == (int)wi + 1) hi = canon.w[j];
        }
        const uint32_t m = (uint32_t)((lo >> sh) | ((hi << 1) << (63 - sh))) & kOwnMidMask;
        const uint32_t y = pairrev32(~m) >> (32 - 2 * kOwnMid);   // reverse complement of the 11 bases
        return owner_fold_mid(m << (32 - 2 * kOwnMid), y << (32 - 2 * kOwnMid));
    }
    uint32_t f = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        f = (f ^ (uint32_t)canon.w[j]) * 0x9E3779B1u;
        f = (f ^ (uint32_t)(canon.w[j] >> 32)) * 0x85EBCA77u;
    }
    return f ^ (f >> 15);
}
__device__ __forceinline__ uint32_t owner_part(uint32_t fold, uint32_t nparts) {
    return __umulhi(fold, nparts);
}

// ---- values derived from the 64-bit hash of the canonical k-mer -------------------------
__device__ __forceinline__ uint64_t hash_sector(uint64_t h, uint32_t sector_shift) {
    return h >> sector_shift;
}
__device__ __forceinline__ uint64_t hash_tag(uint64_t h) {
    return ((h >> 8) & 0xFFFFFFull) << kPosBits;
}
__device__ __forceinline__ uint64_t hash_slot(uint64_t h, uint32_t log2cap) {
    return (h * 0x9E3779B97F4A7C15ull) >> (64 - log2cap);
}

// Bloom bits of a vertex: Q bit positions inside a 32-bit word.  The SAME mask is used in each
// of the 8 edge-slot words of the vertex's sector (word c = in-edge with base c, word 4+c =
// out-edge with base c, canonical orientation): each word is an independent Bloom filter over
// the vertices that have that edge, so sharing the mask between words costs nothing.
__device__ __forceinline__ uint32_t mask_seed(uint64_t h) {  // 32 hash bits the Bloom mask derives from
    return (uint32_t)h ^ ((uint32_t)(h >> 32) * 0x85EBCA77u);
}
// The seed already is 32 mixed hash bits, so the Q bit positions are simply its 5-bit fields (one
// shift + one wrapping shift + a shared OR per field); a second word is derived only for Q > 6.
template <int Q>
__device__ __forceinline__ uint32_t mask_from_seed(uint32_t seed32) {
    uint32_t g = seed32;
    uint32_t m = 0;
#pragma unroll
    for (int t = 0; t < Q; ++t) {
        if (t == 6) g = g * 0x9E3779B1u + 0x7F4A7C15u, g ^= g >> 15;
        m |= 1u << ((g >> (5 * (t % 6))) & 31u);
    }
    return m;
}
__device__ __forceinline__ uint32_t mask_from_seed_rt(uint32_t seed32, uint32_t q) {  // same bits, run-time q
    switch (q) {
        case 1: return mask_from_seed<1>(seed32);
        case 2: return mask_from_seed<2>(seed32);
        case 3: return mask_from_seed<3>(seed32);
        case 4: return mask_from_seed<4>(seed32);
        case 5: return mask_from_seed<5>(seed32);
        case 6: return mask_from_seed<6>(seed32);
        case 7: return mask_from_seed<7>(seed32);
        default: return mask_from_seed<8>(seed32);
    }
}
// Q = 0: the number of bits is a run-time value (kernels of the long k-mers, W > 4, are compiled once for every q)
template <int Q>
__device__ __forceinline__ uint32_t vertex_mask(uint64_t h, uint32_t q) {
    if (Q == 0) return mask_from_seed_rt(mask_seed(h), q);
    return mask_from_seed<(Q ? Q : 1)>(mask_seed(h));
}

// ---- per-thread window over 32 consecutive positions ------------------------------------
// Thread handles positions [32*w, 32*w+32).  After load(): X/Y = k-mer at 32*w; next_feed /
// prev_feed hold the 32 "next" bases (positions 32w+k ..) and the 32 "prev" bases (32w-1 ..);
// valid / prev_n / next_n are per-position bit masks.
template <int W>
struct Window {
    Kmer<W> X, Y;
    uint64_t next_feed, prev_feed;
    uint32_t valid, prev_n, next_n;

    __device__ __forceinline__ void load(const GenomeView& g, uint64_t w, uint32_t k) {
        const uint64_t* __restrict__ codes = g.codes;
        uint64_t cprev = w ? __ldg(codes + w - 1) : 0ull;
        uint64_t c[W + 1];
#pragma unroll
        for (int j = 0; j <= W; ++j) c[j] = __ldg(codes + w + j);
#pragma unroll
        for (int j = 0; j < W; ++j) X.w[j] = c[j];
        X.w[W - 1] &= top_mask<W>(k);
        Y = revcomp<W>(X, k);
        uint32_t kb2 = 2 * (k - 32 * (W - 1));  // 2..62
        next_feed = (c[W - 1] >> kb2) | (c[W] << (64 - kb2));
        prev_feed = (c[0] << 2) | (cprev >> 62);

        uint64_t p0 = w * 32;
        // N bits of positions [p0-1, p0+32+k]
        bool has_n;
        if (w == 0) has_n = true;  // position 0 is always a separator
        else has_n = any_n(g.nmask, p0 - 1, 34 + k);
        if (!has_n) {
            valid = ~0u; prev_n = 0; next_n = 0;
        } else {
            valid = 0; prev_n = 0; next_n = 0;
            uint32_t run = 0;
#pragma unroll 1
            for (uint32_t j = 0; j + 1 < k; ++j) run = load_n(g.nmask, p0 + j) ? 0 : run + 1;
#pragma unroll 1
            for (uint32_t i = 0; i < 32; ++i) {
                run = load_n(g.nmask, p0 + i + k - 1) ? 0 : run + 1;
                if (run >= k) valid |= 1u << i;
                bool pn = (p0 + i == 0) ? true : load_n(g.nmask, p0 + i - 1);
                if (pn) prev_n |= 1u << i;
                if (load_n(g.nmask, p0 + i + k)) next_n |= 1u << i;
            }
        }
    }
};

// neighbours of an occurrence in the orientation of its canonical k-mer
// (candidateoccurence.h:34-47): a = in-edge base, b = out-edge base.
struct Neigh {
    uint32_t a, b;
    bool a_n, b_n;
};
__device__ __forceinline__ Neigh orient(bool fwd_is_canon, uint32_t prv, uint32_t nxt, bool prv_n, bool nxt_n) {
    Neigh r;
    if (fwd_is_canon) { r.a = prv; r.a_n = prv_n; r.b = nxt; r.b_n = nxt_n; }
    else { r.a = 3u - nxt; r.a_n = nxt_n; r.b = 3u - prv; r.b_n = prv_n; }
    return r;
}

// 6-bit occurrence code in canonical orientation: a | a_n << 2 | b << 3 | b_n << 5
__device__ __forceinline__ uint32_t occurrence_code(bool fwd_is_canon, uint32_t prv, uint32_t nxt, uint32_t prv_n, uint32_t nxt_n) {
    const uint32_t f = (prv | (prv_n << 2)) | ((nxt | (nxt_n << 2)) << 3);
    const uint32_t x = f ^ 27u;                       // complement both bases
    const uint32_t r = ((x >> 3) | (x << 3)) & 63u;   // and swap the sides
    return fwd_is_canon ? f : r;
}

// One whole 32-byte filter sector with ONE 256-bit load (sm_100: LDG.E.256): a random sector per
// lane costs one L1 tag cycle per lane and instruction, so halving the instructions doubles the
// rate the L2-resident apply kernels can sustain.
struct Sector {
    uint32_t w[8];
};
__device__ __forceinline__ Sector ld_sector_nc(const uint32_t* p) {  // read-only data (query pass)
    Sector r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ Sector ld_sector_cg(const uint32_t* p) {  // data being updated by atomics (fill pass)
    Sector r;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p) : "memory");
    return r;
}
// word c (0..3) of a half sector without dynamic register indexing
__device__ __forceinline__ uint32_t pick4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t c) {
    uint32_t lo = (c & 1u) ? w1 : w0, hi = (c & 1u) ? w3 : w2;
    return (c & 2u) ? hi : lo;
}

// block-wide sum of a 64-bit value (all threads must call); result valid in thread 0
__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long* smem /* >= 8 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    }
    return t;
}

}  // namespace tpc
