// tpc_bin.cuh -- "binned" filter passes: make the random filter traffic L2-resident.
//
// Measured on B200 (profiles/): uniform random 32-byte sector touches into a 2^36-bit table run
// at ~18.6 G/s (each L2 miss moves a whole 128-byte line from HBM), but at >100 G/s when the
// touched range fits the 126 MB L2.  So instead of touching the filter in genome order
// (k_fill / k_query), the binned path
//   1. k_bin: streams the genome ONCE, computes per owned definite k-mer a 12-byte record
//      {Bloom mask, sector-in-slice | neighbour code, position} and radix-partitions the records
//      by filter slice (slice = 2^slice_log2 bytes, default 64 MiB): CTA-level counting sort in
//      shared memory, coalesced bulk writes per slice (sequential HBM traffic);
//   2. k_apply_fill / k_apply_query: one launch per slice over that slice's records; the slice is
//      fetched from HBM once and every further touch is an L2 hit.
// When one wave of records fits in free HBM the same records serve both passes.
#pragma once
#include "tpc_kernels.cuh"

namespace tpc {

constexpr int kBinHalf = 16;                       // positions per thread per staging round
constexpr int kBinStage = kTileThreads * kBinHalf;  // 4096 records staged per round
constexpr int kBinMaxBuckets = 256;
constexpr int kBinCodeShift = 25;                  // record word 1: sector-in-slice | position bits 32.. | occurrence code << 25
constexpr size_t kBinSmemBytes = kBinMaxBuckets * 8 + kBinMaxBuckets * 4 * 2 + 8 * 4 + kBinStage * 4 * 3;

struct BinView {
    uint32_t* rec;                 // [bucket][3][cap] : mask seed | word1 | relative position
    unsigned long long* count;     // [buckets] records reserved (may exceed cap)
    uint32_t* ov;                  // overflow records {mask seed, word1, relpos, bucket}
    unsigned long long* ov_count;
    uint64_t cap, ov_cap;
    uint32_t bucket_bits;          // log2(#slices)
    uint32_t sib_bits;             // log2(sectors per slice) <= 25
    uint32_t q;                    // Bloom bits per edge (the apply kernels expand the mask seed)
    // Skewed inputs (repeat-rich genomes: every copy of a k-mer goes to one slice): when a slice's array and the overflow
    // list both ran over, the round is re-binned into arrays sized from the exact per-slice counts of the failed attempt:
    // slice b then owns records [off[b], off[b] + capv[b]) of the scratch (three arrays of capv[b] words at rec + 3 off[b]).
    // nullptr = the uniform layout (off[b] = b * cap, capv[b] = cap).
    const unsigned long long* off;
    const unsigned long long* capv;
    __device__ __forceinline__ uint32_t* block_of(uint32_t b) const { return rec + 3 * (off ? off[b] : (uint64_t)b * cap); }
    __device__ __forceinline__ uint64_t cap_of(uint32_t b) const { return capv ? capv[b] : cap; }
};

// record word 1 carries the 6-bit occurrence code in canonical orientation (occurrence_code, tpc_device.cuh)

template <int W>
__global__ void __launch_bounds__(kTileThreads, 2)
k_bin(GenomeView g, KParams kp, BinView bin, uint64_t tile_begin, uint64_t tile_end, uint64_t wave_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* gbase = reinterpret_cast<unsigned long long*>(smem_raw);
    uint32_t* hist = reinterpret_cast<uint32_t*>(gbase + kBinMaxBuckets);
    uint32_t* pref = hist + kBinMaxBuckets;
    uint32_t* warp_tot = pref + kBinMaxBuckets;
    uint32_t* st_a = warp_tot + 8;
    uint32_t* st_b = st_a + kBinStage;
    uint32_t* st_c = st_b + kBinStage;
    const uint32_t nbuckets = 1u << bin.bucket_bits;
    const uint32_t sib_mask = (1u << bin.sib_bits) - 1u;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (uint64_t tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        uint64_t w = tile * kTileThreads + tid;
        Window<W> win;
        win.valid = 0; win.prev_n = 0; win.next_n = 0; win.next_feed = 0; win.prev_feed = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) { win.X.w[j] = 0; win.Y.w[j] = 0; }
        if (w * 32 < g.npos) win.load(g, w, kp.k);
        const bool any_n = (win.prev_n | win.next_n) != 0;
        // k <= 31: both strands of every position by constant shifts out of two code words (and their
        // reverse complement) instead of the rolling dependency chain
        uint64_t c0 = 0, c1 = 0, r_lo = 0, r_hi = 0, kmask = 0;
        if (W == 1 && w * 32 < g.npos) {
            c0 = __ldg(g.codes + w); c1 = __ldg(g.codes + w + 1);
            const uint32_t k2 = 2 * kp.k, s0 = 64 - k2;
            kmask = (~0ull) >> (64 - k2);
            const uint64_t rh = pairrev64(~c0), rl = pairrev64(~c1);
            r_lo = (rl >> s0) | (rh << (64 - s0)); r_hi = rh >> s0;
        }
        const uint64_t rel64 = w * 32 - wave_base;  // position relative to the wave: low 32 bits in word 2,
        const uint32_t relbase = (uint32_t)rel64;   // the bits above in the spare bits of word 1
        const uint32_t relhigh = (uint32_t)(rel64 >> 32) << bin.sib_bits;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            hist[tid] = 0;
            __syncthreads();
            const uint32_t nf32 = half ? (uint32_t)(win.next_feed >> 32) : (uint32_t)win.next_feed;
            const uint32_t pf32 = half ? (uint32_t)(win.prev_feed >> 32) : (uint32_t)win.prev_feed;
            uint32_t rm[kBinHalf], rw[kBinHalf], rk[kBinHalf];
#pragma unroll
            for (int j = 0; j < kBinHalf; ++j) {
                const int i = half * kBinHalf + j;
                const uint32_t nxt = (nf32 >> (2 * j)) & 3u, prv = (pf32 >> (2 * j)) & 3u;
                rk[j] = ~0u; rm[j] = 0; rw[j] = 0;
                if (W == 1) {
                    win.X.w[0] = (i ? ((c0 >> (2 * i)) | (c1 << (64 - 2 * i))) : c0) & kmask;
                    win.Y.w[0] = (i ? ((r_lo >> (64 - 2 * i)) | (r_hi << (2 * i))) : r_hi) & kmask;
                }
                if (win.valid & (1u << i)) {
                    const bool fwd = kmer_less<W>(win.X, win.Y);
                    const Kmer<W> canon = kmer_select<W>(fwd, win.X, win.Y);
                    {   // (k_bin runs unsharded rounds only: every definite k-mer is owned)
                        const uint64_t h = kmer_hash<W>(canon, kp.seed);
                        const uint64_t s = hash_sector(h, kp.sector_shift);
                        uint32_t pn = 0, nn = 0;
                        if (any_n) { pn = (win.prev_n >> i) & 1u; nn = (win.next_n >> i) & 1u; }
                        const uint32_t code = occurrence_code(fwd, prv, nxt, pn, nn);
                        rm[j] = mask_seed(h);
                        rw[j] = ((uint32_t)s & sib_mask) | relhigh | (code << kBinCodeShift);
                        const uint32_t bucket = (uint32_t)(s >> bin.sib_bits);
                        rk[j] = (bucket << 16) | atomicAdd(&hist[bucket], 1u);
                    }
                }
                if (W > 1) roll<W>(win.X, win.Y, nxt, kp.k);
            }
            __syncthreads();
            // exclusive scan of the histogram + global reservation per slice
            uint32_t cnt = hist[tid], incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) warp_tot[wid] = incl;
            __syncthreads();
            uint32_t base = 0;
            for (int j = 0; j < wid; ++j) base += warp_tot[j];
            pref[tid] = base + incl - cnt;
            if (cnt) gbase[tid] = atomicAdd(&bin.count[tid], (unsigned long long)cnt);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kBinHalf; ++j) {
                if (rk[j] != ~0u) {
                    uint32_t idx = pref[rk[j] >> 16] + (rk[j] & 0xFFFFu);
                    st_a[idx] = rm[j];
                    st_b[idx] = rw[j];
                    st_c[idx] = relbase + half * kBinHalf + j;
                }
            }
            __syncthreads();
            // bulk copy-out, one warp per slice: coalesced runs in the slice's record arrays
            for (uint32_t b = wid; b < nbuckets; b += kTileThreads / 32) {
                uint32_t n = hist[b];
                if (!n) continue;
                uint32_t s0 = pref[b];
                unsigned long long gb = gbase[b];
                uint32_t* ra = bin.block_of(b);
                const uint64_t bcap = bin.cap_of(b);
                for (uint32_t j = lane; j < n; j += 32) {
                    unsigned long long dst = gb + j;
                    if (dst < bcap) {
                        __stcs(ra + dst, st_a[s0 + j]);
                        __stcs(ra + bcap + dst, st_b[s0 + j]);
                        __stcs(ra + 2 * bcap + dst, st_c[s0 + j]);
                    } else {
                        unsigned long long o = atomicAdd(bin.ov_count, 1ull);
                        if (o < bin.ov_cap) {
                            uint4 r = make_uint4(st_a[s0 + j], st_b[s0 + j], st_c[s0 + j], b);
                            reinterpret_cast<uint4*>(bin.ov)[o] = r;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}


// ---- sharded variant (nparts > 1): ownership is sparse (1/nparts of the positions).
// k_own decides ownership of EVERY position with as few instructions as possible (the strand-symmetric
// middle-11-mer key of owner_fold, by constant funnel shifts -- no rolling chain, the same cost for
// every k >= 11; high occupancy); k_bin_list compacts the owned positions of a tile into a
// CTA-wide list and runs the expensive part (64-bit hash, record, counting sort by slice) densely
// on it.  One scan serves all the rounds of this GPU: the local round that owns a position
// (1 + round, 0 = none / not a definite k-mer) is written as P bit planes of 1 bit per position.
// ownership of the 32 positions of code word w: pl[j] = bit plane j of the owning local round (1 + part - part_base when that is
// < nlocal, else 0) -- shared by k_own (planes for all the rounds of a call) and the list-driven direct kernels (P = 1, one part)
template <int W, int P>
__device__ __forceinline__ void own_planes_of_word(const GenomeView& g, const KParams& kp, uint32_t part_base, uint32_t nlocal, uint64_t w,
                                                   uint32_t (&pl)[P]) {
#pragma unroll
    for (int j = 0; j < P; ++j) pl[j] = 0;
    if (w * 32 < g.npos) {
        {
            uint32_t valid = ~0u;
            if (w == 0 || any_n(g.nmask, w * 32, 32 + kp.k)) {
                valid = 0;
                uint32_t run = 0;
#pragma unroll 1
                for (uint32_t j = 0; j + 1 < kp.k; ++j) run = load_n(g.nmask, w * 32 + j) ? 0 : run + 1;
#pragma unroll 1
                for (uint32_t i = 0; i < 32; ++i) {
                    run = load_n(g.nmask, w * 32 + i + kp.k - 1) ? 0 : run + 1;
                    if (run >= kp.k) valid |= 1u << i;
                }
            }
            if (kp.k >= kOwnMid) {
                // ownership key = canonical middle 11-mer (owner_fold): A = the 64-base window shifted to the
                // middle of position 0's k-mer, B = its reverse complement; position i's middle 11-mer is bases
                // i.. of A and its reverse complement bases 53-i.. of B -- one constant funnel shift each
                // (any k >= 11, one to four words per k-mer: only the three code words around the middle are read)
                const uint32_t off = (kp.k - kOwnMid) >> 1, off2 = 2 * (off & 31);
                const uint64_t* cw = g.codes + w + (off >> 5);
                const uint64_t c0 = __ldg(cw), c1 = __ldg(cw + 1), c2 = __ldg(cw + 2);
                const uint64_t a0 = off2 ? (c0 >> off2) | (c1 << (64 - off2)) : c0;
                const uint64_t a1 = off2 ? (c1 >> off2) | (c2 << (64 - off2)) : c1;
                const uint64_t b0 = pairrev64(~a1), b1 = pairrev64(~a0);
                const uint32_t A[4] = {(uint32_t)a0, (uint32_t)(a0 >> 32), (uint32_t)a1, (uint32_t)(a1 >> 32)};
                const uint32_t B[5] = {(uint32_t)b0, (uint32_t)(b0 >> 32), (uint32_t)b1, (uint32_t)(b1 >> 32), 0u};
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    constexpr int kTop = 32 - 2 * (int)kOwnMid;
                    const int sa = 2 * i, sb = 2 * (53 - i);
                    const uint32_t m = __funnelshift_r(A[sa >> 5], A[(sa >> 5) + 1], sa & 31) << kTop;
                    const uint32_t y = __funnelshift_r(B[sb >> 5], B[(sb >> 5) + 1], sb & 31) << kTop;
                    const uint32_t local = owner_part(owner_fold_mid(m, y), kp.nparts) - part_base;
                    if (P == 1) pl[0] |= (local < nlocal ? 1u : 0u) << i;
                    else {
                        const uint32_t id = local < nlocal ? local + 1u : 0u;
#pragma unroll
                        for (int j = 0; j < P; ++j) pl[j] |= ((id >> j) & 1u) << i;
                    }
                }
            } else {
            // k < 11 (one word per k-mer): both strands of every position by constant shifts out of two code
            // words and their reverse complement, multiplicative fold of the canonical k-mer
            const uint64_t c0 = __ldg(g.codes + w), c1 = __ldg(g.codes + w + 1);
            const uint32_t k2 = 2 * kp.k;
            const uint64_t kmask = (~0ull) >> (64 - k2);
            const uint64_t rh = pairrev64(~c0), rl = pairrev64(~c1);
            const uint32_t s0 = 64 - k2;
            const uint64_t r_lo = (rl >> s0) | (rh << (64 - s0)), r_hi = rh >> s0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint64_t x = (i ? ((c0 >> (2 * i)) | (c1 << (64 - 2 * i))) : c0) & kmask;
                const uint64_t y = (i ? ((r_lo >> (64 - 2 * i)) | (r_hi << (2 * i))) : r_hi) & kmask;
                Kmer<1> canon;
                canon.w[0] = x < y ? x : y;
                const uint32_t local = owner_part(owner_fold<1>(canon, kp.k), kp.nparts) - part_base;
                if (P == 1) pl[0] |= (local < nlocal ? 1u : 0u) << i;
                else {
                    const uint32_t id = local < nlocal ? local + 1u : 0u;
#pragma unroll
                    for (int j = 0; j < P; ++j) pl[j] |= ((id >> j) & 1u) << i;
                }
            }
            }
#pragma unroll
            for (int j = 0; j < P; ++j) pl[j] &= valid;
        }
    }
}

template <int W, int P>
__global__ void __launch_bounds__(kTileThreads)
k_own(GenomeView g, KParams kp, uint32_t part_base, uint32_t nlocal, uint64_t word_begin, uint64_t word_end, OwnPlanes op) {
    for (uint64_t w = word_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < word_end;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t pl[P];
        own_planes_of_word<W, P>(g, kp, part_base, nlocal, w, pl);
#pragma unroll
        for (int j = 0; j < P; ++j) op.p[j][w] = pl[j];
    }
}

// ---- list-driven direct filter passes (hash-range shards / rounds WITHOUT the binned path: filters of more than 256
// slices, windowed runs).  k_fill / k_query roll the k-mer of EVERY position and test its ownership inline -- with 1/nparts
// of the lanes going on to the hash and the filter sector; at 32 parts (BASELINE config 5) the scan, not the random
// sector traffic, takes the time.  Here ownership is decided per word by the cheap middle-11-mer key (own_planes_of_word,
// no rolling chain), the owned positions of a tile are compacted into a CTA-wide list, and the hash + sector touch run
// densely on it, one list entry per thread, the tile's genome staged in shared memory one tile ahead.
template <int W, int Q, bool QUERY>
__global__ void __launch_bounds__(kTileThreads)
k_direct_list(GenomeView g, uint32_t* __restrict__ filter, KParams kp, uint64_t tile_begin, uint64_t tile_end,
              uint32_t* __restrict__ mask, Counters* ctr, uint32_t* __restrict__ hll) {
    __shared__ TileStage<W> ts;
    __shared__ unsigned long long red[8];
    unsigned long long acc = 0;
    auto own_of = [&](uint64_t t) -> uint32_t {
        uint32_t pl[1];
        own_planes_of_word<W, 1>(g, kp, kp.part, 1u, t * kTileThreads + threadIdx.x, pl);
        return pl[0];
    };
    uint64_t tile = tile_begin + blockIdx.x;
    uint32_t own_next = 0;
    int buf = 0;
    if (tile < tile_end) { own_next = own_of(tile); tile_request(ts, g, tile, 0); }
    for (; tile < tile_end; tile += gridDim.x, buf ^= 1) {
        const uint32_t own = own_next;
        TileGeom tg;
        const uint32_t total = tile_compact(ts, tile, buf, own, tg, [&]() {
            if (tile + gridDim.x < tile_end) { own_next = own_of(tile + gridDim.x); tile_request(ts, g, tile + gridDim.x, buf ^ 1); }
        });
        const uint64_t* s_codes = ts.codes[buf];
        const uint64_t* s_nmask = ts.nmask[buf];
        for (uint32_t e = threadIdx.x; e < total; e += kTileThreads) {
            const uint32_t tp = ts.list[e];
            const uint64_t p = tile * kTilePos + tp;
            const uint32_t lp = tp + tg.c_off, mp = tp + tg.m_off;
            const Kmer<W> X = extract_kmer_smem<W>(s_codes, lp, kp.k);
            const Kmer<W> Y = revcomp<W>(X, kp.k);
            const bool fwd = kmer_less<W>(X, Y);
            const uint64_t h = kmer_hash<W>(kmer_select<W>(fwd, X, Y), kp.seed);
            uint32_t* sec = filter + (hash_sector(h, kp.sector_shift) << 3);
            const uint32_t vm = vertex_mask<Q>(h, kp.q);
            if (!QUERY) {
                const uint32_t code = occurrence_code(fwd, stage_base(s_codes, lp - 1), stage_base(s_codes, lp + kp.k),
                                                      stage_n(s_nmask, mp - 1), stage_n(s_nmask, mp + kp.k));
                acc += fill_vertex(sec, vm, code);
            } else if (query_vertex(sec, vm)) {
                atomicOr(mask + (p >> 5), 1u << (p & 31));
                hll_add(hll, vm, hash_sector(h, kp.sector_shift));
                ++acc;
            }
        }
    }
    unsigned long long t = block_sum(acc, red);
    if (threadIdx.x == 0 && t) atomicAdd(QUERY ? &ctr->marks : &ctr->filter_new, t);
}

constexpr int kBinListMax = kTilePos;
// R = records per thread per staging round (stage = 256 R records); small R = more CTAs per SM
template <int W>
constexpr size_t bin_list_smem_bytes(int R) {
    return kBinMaxBuckets * 8 * 2 + kBinMaxBuckets * 4 * 2 + 8 * 4 + 16 + (size_t)kTileThreads * R * (4 * 3 + 1) + kBinListMax * 2 +
           2 * (TileWords<W>::code + TileWords<W>::mask) * 8;
}

// FUSED: a GPU that runs a single round (4 and more hash-range shards at C3) needs no ownership planes: the ownership word
// of a thread's 32 positions is computed here (own_planes_of_word, the work of k_own) in the issue slots this kernel leaves
// idle while it waits on shared memory and barriers, instead of a separate scan that writes a plane this kernel reads back.
template <int W, int R, bool FUSED>
__global__ void __launch_bounds__(kTileThreads, (R <= 8 ? 4 : 2))
k_bin_list(GenomeView g, KParams kp, BinView bin, uint64_t tile_begin, uint64_t tile_end, uint64_t wave_base, OwnPlanes op) {
    constexpr uint32_t kStage = kTileThreads * R;
    constexpr int kTileCodeWords = TileWords<W>::code, kTileMaskWords = TileWords<W>::mask;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* gbase = reinterpret_cast<unsigned long long*>(smem_raw);    // records reserved before this chunk, per slice
    uint32_t** gptr = reinterpret_cast<uint32_t**>(gbase + kBinMaxBuckets);         // where staged record 0 would go, per slice
    uint64_t* s_stage = reinterpret_cast<uint64_t*>(gptr + kBinMaxBuckets);         // 2 x {code words tw0-1 .., n-mask words (tw0-1)/2 ..}
    uint32_t* hist = reinterpret_cast<uint32_t*>(s_stage + 2 * (kTileCodeWords + kTileMaskWords));
    uint32_t* pref = hist + kBinMaxBuckets;
    uint32_t* warp_tot = pref + kBinMaxBuckets;
    uint32_t* list_total = warp_tot + 8;      // [0] owned positions of the tile, [1] a slice ran over its array in this chunk
    uint32_t* st_a = list_total + 4;
    uint32_t* st_b = st_a + kStage;
    uint32_t* st_c = st_b + kStage;
    uint16_t* list = reinterpret_cast<uint16_t*>(st_c + kStage);
    uint8_t* st_k = reinterpret_cast<uint8_t*>(list + kBinListMax);           // slice of each staged record
    const uint32_t sib_mask = (1u << bin.sib_bits) - 1u;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t kmask1 = top_mask<1>(kp.k < 32 ? kp.k : 31);              // W == 1 only
    const uint32_t nxt_shift = 2 * kp.k + 2;                                  // W == 1, k <= 29: next base inside the 32-base window

    // Software pipeline over the CTA's tiles: the ownership word and the packed-genome words of tile t+1
    // are requested (register / cp.async into the other staging buffer) before tile t is processed, so no
    // HBM latency sits between the barriers of a tile.
    auto request_tile = [&](uint64_t t, int buf) {
        const uint64_t w0 = t * kTileThreads, cb = w0 ? w0 - 1 : 0, mb = cb >> 1;
        uint64_t* sc = s_stage + buf * (kTileCodeWords + kTileMaskWords);
        uint64_t* sm = sc + kTileCodeWords;
        for (int j = tid; j < kTileCodeWords; j += kTileThreads) tile_cp_async8(sc + j, g.codes + cb + j);
        for (int j = tid; j < kTileMaskWords; j += kTileThreads) tile_cp_async8(sm + j, g.nmask + mb + j);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto own_of = [&](uint64_t w) -> uint32_t {
        if (FUSED) {
            uint32_t pl[1];
            own_planes_of_word<W, 1>(g, kp, kp.part, 1u, w, pl);
            return pl[0];
        }
        return own_word(op, w);
    };
    uint64_t tile = tile_begin + blockIdx.x;
    uint32_t own_next = 0;
    int buf = 0;
    if (tile < tile_end) {
        own_next = own_of(tile * kTileThreads + tid);
        request_tile(tile, 0);
    }
    for (; tile < tile_end; tile += gridDim.x, buf ^= 1) {
        uint32_t own = own_next;
        const uint64_t* s_codes = s_stage + buf * (kTileCodeWords + kTileMaskWords);
        const uint64_t* s_nmask = s_codes + kTileCodeWords;
        const uint64_t tw0 = tile * kTileThreads;              // first code word of the tile
        const uint64_t cw_base = tw0 ? tw0 - 1 : 0;            // s_codes[0] = word cw_base
        const uint64_t mw_base = cw_base >> 1;                 // s_nmask[0] = word mw_base
        // CTA-wide list of owned positions (tile-relative, in position order)
        uint32_t cnt = __popc(own), incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // this tile's genome words (my copies) have landed
        __syncthreads();  // ... everybody's; the previous tile's readers of list / warp_tot / the other buffer are done
        if (lane == 31) warp_tot[wid] = incl;
        if (tile + gridDim.x < tile_end) {                     // next tile: in flight during this tile's dense phase
            own_next = own_of((tile + gridDim.x) * kTileThreads + tid);
            request_tile(tile + gridDim.x, buf ^ 1);
        }
        uint64_t n_any = 0;
        for (int j = tid; j < kTileMaskWords; j += kTileThreads) n_any |= s_nmask[j];
        // no 'N' anywhere near the tile (all but the tiles holding a record boundary or an N run): the
        // dense phase then needs no n-mask look-ups at all
        const bool tile_has_n = __syncthreads_or(n_any != 0) != 0;
        uint32_t off = incl - cnt;
        for (int j = 0; j < wid; ++j) off += warp_tot[j];
        if (tid == kTileThreads - 1) list_total[0] = off + cnt;
        while (own) {
            int i = __ffs(own) - 1;
            own &= own - 1;
            list[off++] = (uint16_t)(tid * 32 + i);
        }
        __syncthreads();
        const uint32_t total = list_total[0];
        const uint32_t c_off = (uint32_t)(tw0 - cw_base) * 32;          // tile-local -> s_codes-local position
        const uint32_t m_off = (uint32_t)(tw0 * 32 - mw_base * 64);     // tile-local -> s_nmask-local position
        const uint64_t tile_rel = tile * kTilePos - wave_base;
        const uint32_t rel_high = (uint32_t)(tile_rel >> 32) << bin.sib_bits;   // a tile never straddles a 2^32 boundary of the wave
        const uint32_t rel_low = (uint32_t)tile_rel;

        // dense processing of the list, kStage records per round
        for (uint32_t c0 = 0; c0 < total; c0 += kStage) {
            const uint32_t n = min(kStage, total - c0);
            hist[tid] = 0;
            if (tid == 0) list_total[1] = 0;
            __syncthreads();
            uint32_t rm[R], rw[R], rk[R];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const uint32_t e = tid + j * kTileThreads;
                rk[j] = ~0u; rm[j] = 0; rw[j] = 0;
                if (e < n) {
                    const uint32_t tp = list[c0 + e];                // tile-local position (>= 1 when tile 0: position 0 is N)
                    const uint32_t lp = tp + c_off;
                    Kmer<W> X, Y;
                    uint32_t prv, nxt;
                    if (W == 1) {
                        // one unaligned 32-base window starting at the previous base holds prev | k-mer | next (k <= 29)
                        const uint32_t q0 = lp - 1, wi = q0 >> 5, sh = 2 * (q0 & 31);
                        const uint64_t lo = s_codes[wi], hi = s_codes[wi + 1];
                        const uint64_t win = (lo >> sh) | ((hi << 1) << (63 - sh));
                        prv = (uint32_t)win & 3u;
                        X.w[0] = (win >> 2) & kmask1;
                        nxt = (kp.k >= 31 ? (uint32_t)(hi >> sh) : (uint32_t)(win >> nxt_shift)) & 3u;   // k = 31: base 32 of the window
                    } else {
                        X = extract_kmer_smem<W>(s_codes, lp, kp.k);
                        const uint32_t pp = lp - 1, np = lp + kp.k;
                        prv = (uint32_t)(s_codes[pp >> 5] >> (2 * (pp & 31))) & 3u;
                        nxt = (uint32_t)(s_codes[np >> 5] >> (2 * (np & 31))) & 3u;
                    }
                    Y = revcomp<W>(X, kp.k);
                    const bool fwd = kmer_less<W>(X, Y);
                    const uint64_t h = kmer_hash<W>(kmer_select<W>(fwd, X, Y), kp.seed);
                    uint32_t pn = 0, nn = 0;
                    if (tile_has_n) {
                        const uint32_t pm = tp + m_off - 1, nm = tp + m_off + kp.k;
                        pn = (uint32_t)(s_nmask[pm >> 6] >> (pm & 63)) & 1u;
                        nn = (uint32_t)(s_nmask[nm >> 6] >> (nm & 63)) & 1u;
                    }
                    const uint32_t code = occurrence_code(fwd, prv, nxt, pn, nn);
                    const uint64_t s = hash_sector(h, kp.sector_shift);
                    rm[j] = mask_seed(h);
                    rw[j] = ((uint32_t)s & sib_mask) | rel_high | (code << kBinCodeShift);
                    const uint32_t bucket = (uint32_t)(s >> bin.sib_bits);
                    rk[j] = (bucket << 16) | atomicAdd(&hist[bucket], 1u);
                }
            }
            __syncthreads();
            uint32_t hc = hist[tid], hincl = hc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, hincl, o);
                if (lane >= o) hincl += v;
            }
            if (lane == 31) warp_tot[wid] = hincl;
            __syncthreads();
            uint32_t base = 0;
            for (int j = 0; j < wid; ++j) base += warp_tot[j];
            const uint32_t my_pref = base + hincl - hc;
            pref[tid] = my_pref;
            // thread tid reserves the run of slice tid; the answer is only needed after the scatter below
            unsigned long long gb = 0;
            if (hc) gb = atomicAdd(&bin.count[tid], (unsigned long long)hc);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (rk[j] != ~0u) {
                    uint32_t idx = pref[rk[j] >> 16] + (rk[j] & 0xFFFFu);
                    st_a[idx] = rm[j];
                    st_b[idx] = rw[j];
                    st_c[idx] = rel_low + list[c0 + tid + j * kTileThreads];
                    st_k[idx] = (uint8_t)(rk[j] >> 16);
                }
            }
            if (hc) {
                gbase[tid] = gb;
                gptr[tid] = bin.block_of(tid) + gb - my_pref;   // + stage index = the record's place
                if (gb + hc > bin.cap_of(tid) || bin.capv) list_total[1] = 1;   // (per-slice capacities: the general copy-out)
            }
            __syncthreads();
            // copy-out: one thread per staged record (records of a slice are contiguous in the stage, so
            // neighbouring threads write neighbouring words of the slice's arrays)
            if (list_total[1] == 0) {
                const uint64_t cap = bin.cap;
                for (uint32_t idx = tid; idx < n; idx += kTileThreads) {
                    uint32_t* ra = gptr[st_k[idx]] + idx;
                    __stcs(ra, st_a[idx]);
                    __stcs(ra + cap, st_b[idx]);
                    __stcs(ra + 2 * cap, st_c[idx]);
                }
            } else {   // a slice's array is full (repeat-rich input): per-record bound check, overflow list
                for (uint32_t idx = tid; idx < n; idx += kTileThreads) {
                    const uint32_t b = st_k[idx];
                    const unsigned long long dst = gbase[b] + (idx - pref[b]);
                    const uint64_t bcap = bin.cap_of(b);
                    if (dst < bcap) {
                        uint32_t* ra = bin.block_of(b) + dst;
                        __stcs(ra, st_a[idx]);
                        __stcs(ra + bcap, st_b[idx]);
                        __stcs(ra + 2 * bcap, st_c[idx]);
                    } else {
                        unsigned long long o = atomicAdd(bin.ov_count, 1ull);
                        if (o < bin.ov_cap) reinterpret_cast<uint4*>(bin.ov)[o] = make_uint4(st_a[idx], st_b[idx], st_c[idx], b);
                    }
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace tpc
