// tpc_kernels.cuh -- the hand-written sm_100a kernels of the junction-finding path.
//
//   k_fill      pass 1a  FilterFillerWorker        (vertexenumerator.h:995-1105)
//   k_query     pass 1b  CandidateCheckingWorker   (vertexenumerator.h:586-704)
//   k_insert    pass 2   CandidateFinalFilteringWorker (vertexenumerator.h:708-829)
//   k_classify           TrueBifurcations          (vertexenumerator.h:1228-1256)
//   k_build_index        BifurcationStorage::Init  (bifurcationstorage.h:27-66)
//   k_ends / k_emit_count / k_emit_write
//                        EdgeConstructionWorker    (vertexenumerator.h:856-993)
//                        + JunctionPositionWriter  (junctionapi.h:107-137)
//
// Filter layout ("vertex-blocked"): the 2^f-bit filter is an array of 32-byte sectors.  A
// canonical k-mer (vertex) owns one sector chosen by its hash; the 8 32-bit words of the
// sector are the Bloom words of its 8 possible incident edges (word c = in-edge with base c,
// word 4+c = out-edge with base c, canonical orientation); an edge sets Q bits in its word.
// An edge of the de Bruijn graph is therefore recorded twice (once per endpoint) but both the
// fill and the 8-edge query of a k-mer touch exactly ONE 32-byte sector.  The reference hashes
// every (k+1)-mer independently (vertexrollinghash.h:144-252): q sectors per insert, 6..8q per
// query.  The bit layout of the filter is unobservable (hash seeds are random in the
// reference); only "no false negatives" matters, which holds by construction.
#pragma once
#include "tpc_device.cuh"
#include "tpc_tile.cuh"

namespace tpc {

struct Counters {
    unsigned long long filter_new;   // fill: (vertex, slot) items that had to set bits
    unsigned long long marks;        // query: candidate marks
    unsigned long long distinct;     // insert: slots claimed
    unsigned long long overflow;     // insert: probe sequence exhausted
    unsigned long long junctions;    // classify: junctions appended
    unsigned long long dropped;      // classify: dropped by abundance
    unsigned long long list_incomplete;  // insert from the mark list: regions that had overflowed (pass redone from the mask)
};

// HyperLogLog sketch (4096 registers) of the candidate k-mers, updated by the query kernels; the
// host sizes the exact candidate table from it (the reference grows std::unordered_set instead).
constexpr int kHllBits = 12;
__device__ __forceinline__ void hll_add(uint32_t* __restrict__ hll, uint32_t vertex_mask_bits, uint64_t sector) {
    uint64_t hh = fmix64(((uint64_t)vertex_mask_bits << 32) ^ sector);
    uint32_t idx = (uint32_t)hh & ((1u << kHllBits) - 1u);
    uint32_t rho = (uint32_t)__clzll((hh >> kHllBits) << kHllBits | (1ull << (kHllBits - 1))) + 1u;
    if (rho > __ldcg(hll + idx)) atomicMax(hll + idx, rho);
}

// ------------------------------------------------------------------------------------------
// pass 1a: fill
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t filter_set(uint32_t* word, uint32_t cur, uint32_t m) {
    if ((cur & m) != m) {
        atomicOr(word, m);
        return 1;
    }
    return 0;
}

// Record the in-edge and the out-edge of one occurrence in the vertex's sector (already loaded with one
// 256-bit load): test, then atomicOr only on the words that miss bits.  `code` is the occurrence code
// in canonical orientation (occurrence_code).  An 'N' neighbour is unique: make every occurrence of
// this k-mer a candidate by recording two distinct dummy edges (h:1044-1058).
__device__ __forceinline__ uint32_t fill_sector(uint32_t* sec, const Sector& s, uint32_t m, uint32_t code) {
    const uint32_t a = code & 3u, b = (code >> 3) & 3u;
    uint32_t fresh = 0;
    if ((code & 0x24u) == 0) {  // no 'N' neighbour (all but a few occurrences)
        fresh += filter_set(sec + a, pick4(s.w[0], s.w[1], s.w[2], s.w[3], a), m);
        fresh += filter_set(sec + 4 + b, pick4(s.w[4], s.w[5], s.w[6], s.w[7], b), m);
        return fresh;
    }
    if (!(code & 4u)) fresh += filter_set(sec + a, pick4(s.w[0], s.w[1], s.w[2], s.w[3], a), m);
    else { fresh += filter_set(sec + 0, s.w[0], m); fresh += filter_set(sec + 3, s.w[3], m); }
    if (!(code & 32u)) fresh += filter_set(sec + 4 + b, pick4(s.w[4], s.w[5], s.w[6], s.w[7], b), m);
    else { fresh += filter_set(sec + 4, s.w[4], m); fresh += filter_set(sec + 7, s.w[7], m); }
    return fresh;
}
__device__ __forceinline__ uint32_t fill_vertex(uint32_t* sec, uint32_t m, uint32_t code) {
    Sector s = ld_sector_cg(sec);
    return fill_sector(sec, s, m, code);
}

// All 8 edge queries of one k-mer from its sector (h:640-660).  The fill pass has recorded the edges
// of THIS occurrence (or two dummy edges per 'N' side) in the same sector, so "the edge present here
// counts once, any other recorded edge counts too" is simply: two or more of the four in-edge words,
// or of the four out-edge words, hold the vertex's mask.  No neighbour information is needed.
__device__ __forceinline__ bool query_sector(const Sector& s, uint32_t m) {
    uint32_t in_cnt = 0, out_cnt = 0;
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
        in_cnt += ((s.w[c] & m) == m) ? 1u : 0u;
        out_cnt += ((s.w[4 + c] & m) == m) ? 1u : 0u;
    }
    return in_cnt > 1 || out_cnt > 1;
}
__device__ __forceinline__ bool query_vertex(const uint32_t* sec, uint32_t m) {
    Sector s = ld_sector_nc(sec);
    return query_sector(s, m);
}

template <int W, int Q>
__global__ void __launch_bounds__(kTileThreads)
k_fill(GenomeView g, uint32_t* __restrict__ filter, KParams kp, uint64_t tile_begin, uint64_t ntiles, Counters* ctr) {
    __shared__ unsigned long long red[8];
    unsigned long long fresh = 0;
    for (uint64_t tile = tile_begin + blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint64_t w = tile * kTileThreads + threadIdx.x;
        if (w * 32 >= g.npos) continue;
        Window<W> win;
        win.load(g, w, kp.k);
        if (win.valid == 0) continue;
        uint64_t nf = win.next_feed, pf = win.prev_feed;
#pragma unroll 2
        for (int i = 0; i < 32; ++i) {
            uint32_t nxt = (uint32_t)nf & 3u, prv = (uint32_t)pf & 3u;
            nf >>= 2; pf >>= 2;
            if ((win.valid >> i) & 1u) {
                bool fwd = kmer_less<W>(win.X, win.Y);
                Kmer<W> canon = kmer_select<W>(fwd, win.X, win.Y);
                if (kp.nparts == 1 || owner_part(owner_fold<W>(canon, kp.k), kp.nparts) == kp.part) {
                    uint64_t h = kmer_hash<W>(canon, kp.seed);
                    uint32_t code = occurrence_code(fwd, prv, nxt, (win.prev_n >> i) & 1u, (win.next_n >> i) & 1u);
                    uint32_t* sec = filter + (hash_sector(h, kp.sector_shift) << 3);
                    fresh += fill_vertex(sec, vertex_mask<Q>(h, kp.q), code);
                }
            }
            roll<W>(win.X, win.Y, nxt, kp.k);
        }
    }
    unsigned long long t = block_sum(fresh, red);
    if (threadIdx.x == 0 && t) atomicAdd(&ctr->filter_new, t);
}

// ------------------------------------------------------------------------------------------
// pass 1b: query the 4 in-edges and 4 out-edges of every owned k-mer -> candidate mask
// ------------------------------------------------------------------------------------------
template <int W, int Q>
__global__ void __launch_bounds__(kTileThreads)
k_query(GenomeView g, const uint32_t* __restrict__ filter, KParams kp, uint64_t tile_begin, uint64_t ntiles,
        uint32_t* __restrict__ mask, int accumulate, Counters* ctr, uint32_t* __restrict__ hll) {
    __shared__ unsigned long long red[8];
    unsigned long long marks = 0;
    for (uint64_t tile = tile_begin + blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint64_t w = tile * kTileThreads + threadIdx.x;
        if (w * 32 >= g.npos) continue;
        Window<W> win;
        win.load(g, w, kp.k);
        uint32_t out = 0;
        if (win.valid != 0) {
            uint64_t nf = win.next_feed;
#pragma unroll 2
            for (int i = 0; i < 32; ++i) {
                uint32_t nxt = (uint32_t)nf & 3u;
                nf >>= 2;
                if ((win.valid >> i) & 1u) {
                    bool fwd = kmer_less<W>(win.X, win.Y);
                    Kmer<W> canon = kmer_select<W>(fwd, win.X, win.Y);
                    if (kp.nparts == 1 || owner_part(owner_fold<W>(canon, kp.k), kp.nparts) == kp.part) {
                        uint64_t h = kmer_hash<W>(canon, kp.seed);
                        const uint32_t* sec = filter + (hash_sector(h, kp.sector_shift) << 3);
                        uint32_t vm = vertex_mask<Q>(h, kp.q);
                        if (query_vertex(sec, vm)) {
                            out |= 1u << i;
                            hll_add(hll, vm, hash_sector(h, kp.sector_shift));
                        }
                    }
                }
                roll<W>(win.X, win.Y, nxt, kp.k);
            }
        }
        if (accumulate) { if (out) mask[w] |= out; }
        else mask[w] = out;
        marks += __popc(out);
    }
    unsigned long long t = block_sum(marks, red);
    if (threadIdx.x == 0 && t) atomicAdd(&ctr->marks, t);
}

// ------------------------------------------------------------------------------------------
// pass 2: exact hash set of the candidates
// ------------------------------------------------------------------------------------------
// One occurrence at position p: everything the table operations need.
template <int W>
struct Occ {
    Kmer<W> X, Y;
    uint64_t h;
    uint32_t fold;  // owner_fold of the canonical k-mer
    bool fwd;       // X is the canonical strand
};

template <int W>
__device__ __forceinline__ Occ<W> occurrence_at(const GenomeView& g, uint64_t p, const KParams& kp) {
    Occ<W> o;
    o.X = extract_kmer<W>(g.codes, p, kp.k);
    o.Y = revcomp<W>(o.X, kp.k);
    o.fwd = kmer_less<W>(o.X, o.Y);
    Kmer<W> canon = kmer_select<W>(o.fwd, o.X, o.Y);
    o.fold = owner_fold<W>(canon, kp.k);
    o.h = kmer_hash<W>(canon, kp.seed);
    return o;
}

// does the slot's representative occurrence spell the same k-mer (either strand)?
// returns 0 = no, 1 = same strand as X, 2 = opposite strand
template <int W>
__device__ __forceinline__ int match_rep(const GenomeView& g, unsigned long long rep, const Occ<W>& o, uint32_t k) {
    Kmer<W> r = extract_kmer<W>(g.codes, rep & kPosMask, k);
    if (kmer_eq<W>(r, o.X)) return 1;
    if (kmer_eq<W>(r, o.Y)) return 2;
    return 0;
}

// One marked occurrence (k-mer X at position p, neighbours prv / nxt, 'N' flags) into the candidate table:
// find or claim the slot of its canonical k-mer, keep the smallest position, OR the neighbour sets in canonical
// orientation (candidateoccurence.h:25-50; h:778-796).  k <= 31: key inline in the slot (CAS on the key); else
// position-identified slot (CAS on tag|position, atomicMin).  Returns false when the mark is not this round's.
template <int W>
__device__ __forceinline__ bool insert_occurrence(const GenomeView& g, const KParams& kp, const TableView& T, Counters* ctr, uint64_t p,
                                                  const Kmer<W>& X, uint32_t prv, uint32_t nxt, bool prv_n, bool nxt_n, bool check_owner,
                                                  unsigned long long& claimed) {
    const uint64_t capmask = (1ull << T.log2cap) - 1;
    const uint64_t probe_limit = capmask < 8192 ? capmask : 8192;  // a longer run means the table is too full: host grows it
    Occ<W> o;
    o.X = X;
    o.Y = revcomp<W>(o.X, kp.k);
    o.fwd = kmer_less<W>(o.X, o.Y);
    const Kmer<W> canon = kmer_select<W>(o.fwd, o.X, o.Y);
    if (check_owner && owner_part(owner_fold<W>(canon, kp.k), kp.nparts) != kp.part) return false;  // marked in another round
    o.h = kmer_hash<W>(canon, kp.seed);
    Neigh nb = orient(o.fwd, prv, nxt, prv_n, nxt_n);
    unsigned long long want = 0;
    if (!nb.a_n) want |= 1ull << nb.a;
    if (!nb.b_n) want |= 16ull << nb.b;
    uint64_t idx = hash_slot(o.h, T.log2cap);
    Slot* s = nullptr;
    unsigned long long meta = 0;
    if (W == 1 && T.inline_keys) {
        const unsigned long long key1 = canon.w[0] + 1ull;
        for (uint64_t probe = 0; probe <= probe_limit; ++probe, idx = (idx + 1) & capmask) {
            Slot* cand = T.slots + idx;
            ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(cand));
            if (v.x == 0) {
                v.x = atomicCAS(&cand->rep, 0ull, key1);
                if (v.x == 0) { s = cand; ++claimed; v.x = key1; v.y = 0; }
            }
            if (v.x == key1) { s = cand; meta = v.y; break; }
        }
        if (!s) { atomicAdd(&ctr->overflow, 1ull); return true; }
        // flags by OR, first position (with the strand of its occurrence) by a CAS-min on the high bits (0 = not yet set)
        if ((meta & want) != want) meta = atomicOr(&s->meta, want) | want;
        const unsigned long long field = ((unsigned long long)p << 1) | (o.fwd ? 1ull : 0ull);
        while ((meta >> kInlinePosShift) == 0 || (meta >> kInlinePosShift) > field) {
            unsigned long long neu = (meta & ((1ull << kInlinePosShift) - 1)) | (field << kInlinePosShift);
            unsigned long long old = atomicCAS(&s->meta, meta, neu);
            if (old == meta) break;
            meta = old;
        }
    } else {
        unsigned long long mine = hash_tag(o.h) | p;
        for (uint64_t probe = 0; probe <= probe_limit; ++probe, idx = (idx + 1) & capmask) {
            Slot* cand = T.slots + idx;
            ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(cand));
            unsigned long long rep = v.x;
            if (rep == 0) {
                rep = atomicCAS(&cand->rep, 0ull, mine);
                if (rep == 0) { s = cand; ++claimed; break; }
                v.y = 0;
            }
            if ((rep >> kPosBits) == (mine >> kPosBits) && match_rep<W>(g, rep, o, kp.k)) {
                s = cand; meta = v.y;
                if (mine < rep) atomicMin(&cand->rep, mine);  // keep the first occurrence
                break;
            }
        }
        if (!s) { atomicAdd(&ctr->overflow, 1ull); return true; }
        if ((meta & want) != want) atomicOr(&s->meta, want);
        if (kp.count_occurrences) atomicAdd(&s->meta, 1ull << kMetaCountShift);
    }
    if (nb.a_n) { if (atomicOr(&s->meta, kMetaInN1) & kMetaInN1) atomicOr(&s->meta, kMetaInN2); }
    if (nb.b_n) { if (atomicOr(&s->meta, kMetaOutN1) & kMetaOutN1) atomicOr(&s->meta, kMetaOutN2); }
    return true;
}

// Candidate marks are sparse (a few % of the positions), so the marks of a tile are compacted into a
// CTA-wide list and inserted one per thread (tpc_tile.cuh).  `op` (optional) selects this round's
// marks through the ownership planes; without planes ownership is recomputed from the k-mer.
template <int W>
__global__ void __launch_bounds__(kTileThreads)
k_insert(GenomeView g, const uint32_t* __restrict__ mask, KParams kp, uint64_t tile_begin, uint64_t ntiles, TableView T, Counters* ctr, OwnPlanes op) {
    __shared__ TileStage<W> ts;
    unsigned long long claimed = 0;
    // this round's marks of a tile (software-pipelined: the next tile's words are requested while this one is processed)
    auto marks_of = [&](uint64_t t) -> uint32_t {
        const uint64_t w = t * kTileThreads + threadIdx.x;
        uint32_t m = __ldcs(mask + w);
        if (op.n) m &= own_word(op, w);
        return m;
    };
    uint64_t tile = tile_begin + blockIdx.x;
    uint32_t m_next = 0;
    int buf = 0;
    if (tile < ntiles) { m_next = marks_of(tile); tile_request(ts, g, tile, 0); }
    for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const uint32_t m = m_next;
        TileGeom tg;
        const uint32_t total = tile_compact(ts, tile, buf, m, tg, [&]() {
            if (tile + gridDim.x < ntiles) { m_next = marks_of(tile + gridDim.x); tile_request(ts, g, tile + gridDim.x, buf ^ 1); }
        });
        const uint64_t* s_codes = ts.codes[buf];
        const uint64_t* s_nmask = ts.nmask[buf];
        for (uint32_t e = threadIdx.x; e < total; e += kTileThreads) {
            const uint32_t tp = ts.list[e];
            const uint64_t p = tile * kTilePos + tp;
            const uint32_t lp = tp + tg.c_off, mp = tp + tg.m_off;
            const Kmer<W> X = extract_kmer_smem<W>(s_codes, lp, kp.k);
            insert_occurrence<W>(g, kp, T, ctr, p, X, stage_base(s_codes, lp - 1), stage_base(s_codes, lp + kp.k),
                                 stage_n(s_nmask, mp - 1) != 0, stage_n(s_nmask, mp + kp.k) != 0, !op.n && kp.nparts > 1, claimed);
        }
    }
    if (claimed) atomicAdd(&ctr->distinct, claimed);
}

// The same from the mark list the binned query kernels append to (MarkList, tpc_kernels_common.cuh): when the marks are
// sparse (hash-range shards: 1/N of them per GPU) walking every mask word and staging every tile costs more than the
// inserts; here one thread takes one listed position and reads its k-mer straight from the packed genome.
struct MarkList {
    unsigned long long* entries;     // [regions][region_cap] positions; region = CTA index of the appending kernel
    uint32_t* counts;                // [regions] entries appended (may exceed region_cap: the list is then incomplete)
    uint32_t regions;
    uint32_t region_cap;
};

template <int W>
__global__ void __launch_bounds__(256)
k_insert_list(GenomeView g, MarkList ml, KParams kp, TableView T, Counters* ctr) {
    unsigned long long claimed = 0;
    for (uint32_t r = blockIdx.x; r < ml.regions; r += gridDim.x) {
        uint32_t n = ml.counts[r];
        if (n > ml.region_cap) {   // (the host then redoes the pass from the mask)
            if (threadIdx.x == 0) atomicAdd(&ctr->list_incomplete, 1ull);
            n = ml.region_cap;
        }
        const unsigned long long* e = ml.entries + (uint64_t)r * ml.region_cap;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t p = __ldcs(e + i);
            const Kmer<W> X = extract_kmer<W>(g.codes, p, kp.k);
            insert_occurrence<W>(g, kp, T, ctr, p, X, load_base(g.codes, p - 1), load_base(g.codes, p + kp.k),
                                 load_n(g.nmask, p - 1) != 0, load_n(g.nmask, p + kp.k) != 0, false, claimed);
        }
    }
    if (claimed) atomicAdd(&ctr->distinct, claimed);
}

// ------------------------------------------------------------------------------------------
// junction index J: sorted first-occurrence positions -> ids 1..J (bifurcationstorage.h:27-66)
// ------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256)
k_build_index(GenomeView g, const unsigned long long* __restrict__ sorted_pos, uint64_t n, KParams kp, TableView J) {
    const uint64_t capmask = (1ull << J.log2cap) - 1;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p = sorted_pos[i] & kPosMask;
        Occ<W> o = occurrence_at<W>(g, p, kp);
        unsigned long long mine = hash_tag(o.h) | p;
        if (W == 1 && J.inline_keys) mine = ((o.fwd ? o.X.w[0] : o.Y.w[0]) + 1ull) | (o.fwd ? (1ull << 63) : 0ull);
        uint64_t idx = hash_slot(o.h, J.log2cap);
        for (;; idx = (idx + 1) & capmask) {
            if (atomicCAS(&J.slots[idx].rep, 0ull, mine) == 0ull) {
                J.slots[idx].meta = i + 1;
                break;
            }
        }
    }
}

// every position that starts a definite k-mer (emit without a candidate mask: the windowed runs test all of them
// against the junction index)
template <int W>
__global__ void __launch_bounds__(kTileThreads)
k_valid_mask(GenomeView g, KParams kp, uint64_t word_begin, uint64_t word_end, uint32_t* __restrict__ mask) {
    for (uint64_t w = word_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < word_end; w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t valid = 0;
        if (w * 32 < g.npos) {
            valid = ~0u;
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
        }
        mask[w] = valid;
    }
}

// -> signed id (+ same strand as the first occurrence, - opposite; bifurcationstorage.h:100-128),
// 0 when the k-mer is not a junction
template <int W>
__device__ __forceinline__ long long lookup_id(const GenomeView& g, const TableView& J, const Occ<W>& o, uint32_t k) {
    const uint64_t capmask = (1ull << J.log2cap) - 1;
    if (W == 1 && J.inline_keys) {
        const unsigned long long key1 = (o.fwd ? o.X.w[0] : o.Y.w[0]) + 1ull;
        for (uint64_t idx = hash_slot(o.h, J.log2cap);; idx = (idx + 1) & capmask) {
            ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(J.slots + idx));
            if (v.x == 0) return 0;
            if ((v.x & ~(1ull << 63)) == key1) return ((v.x >> 63) != 0) == o.fwd ? (long long)v.y : -(long long)v.y;
        }
    }
    unsigned long long tag = hash_tag(o.h) >> kPosBits;
    for (uint64_t idx = hash_slot(o.h, J.log2cap);; idx = (idx + 1) & capmask) {
        ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(J.slots + idx));
        if (v.x == 0) return 0;
        if ((v.x >> kPosBits) == tag) {
            int mr = match_rep<W>(g, v.x, o, k);
            if (mr) return mr == 1 ? (long long)v.y : -(long long)v.y;
        }
    }
}

struct RecordTable {
    const uint64_t* __restrict__ start;   // first position of each record
    const uint64_t* __restrict__ len;
    const uint32_t* __restrict__ sep_before;  // separators written before the first record of the sequence
    uint64_t n;
};

// Sequence ends are always emitted (h:942-948): mark those that are not junction occurrences.
template <int W>
__global__ void __launch_bounds__(256)
k_ends(GenomeView g, RecordTable rt, KParams kp, TableView J, uint32_t* __restrict__ stubmask,
       uint64_t pos_begin, uint64_t pos_end) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < rt.n; r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t len = rt.len[r];
        if (len < kp.k) continue;  // h:1177: no task, no records
        uint64_t ends[2] = {rt.start[r], rt.start[r] + len - kp.k};
        int n_ends = ends[0] == ends[1] ? 1 : 2;
        for (int e = 0; e < n_ends; ++e) {
            uint64_t p = ends[e];
            if (p < pos_begin || p >= pos_end) continue;
            bool junction = false;
            if (!any_n(g.nmask, p, kp.k)) {
                Occ<W> o = occurrence_at<W>(g, p, kp);
                junction = lookup_id<W>(g, J, o, kp.k) != 0;
            }
            if (!junction) atomicOr(stubmask + (p >> 5), 1u << (p & 31));
        }
    }
}

// Resolve the candidate mask against the junction index (clears Bloom false positives in
// place) and count records / stubs per tile.
template <int W>
__global__ void __launch_bounds__(kTileThreads)
k_emit_count(GenomeView g, uint32_t* __restrict__ mask, const uint32_t* __restrict__ stubmask, KParams kp,
             TableView J, uint64_t tile_begin, uint64_t tile_end,
             unsigned long long* __restrict__ tile_records, unsigned long long* __restrict__ tile_stubs) {
    __shared__ unsigned long long red[8];
    for (uint64_t tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        uint64_t w = tile * kTileThreads + threadIdx.x;
        uint32_t keep = 0, stub = 0;
        if (w * 32 < g.npos) {
            uint32_t m = mask[w];
            stub = stubmask[w];
            uint32_t todo = m & ~stub;
            while (todo) {
                int i = __ffs(todo) - 1;
                todo &= todo - 1;
                Occ<W> o = occurrence_at<W>(g, w * 32 + i, kp);
                if (lookup_id<W>(g, J, o, kp.k) != 0) keep |= 1u << i;
            }
            if (keep != m) mask[w] = keep;
        }
        unsigned long long packed = ((unsigned long long)__popc(stub) << 32) | (unsigned)(__popc(keep) + __popc(stub));
        unsigned long long t = block_sum(packed, red);
        if (threadIdx.x == 0) {
            tile_records[tile - tile_begin] = t & 0xFFFFFFFFull;
            tile_stubs[tile - tile_begin] = t >> 32;
        }
    }
}

// Ordered emission of the 12-byte records {u32 pos, i64 id} (junctionapi.h:118-132).
// Unit index of a record = (records before it) + (index of its sequence); the separators that
// precede the first record of a sequence are written by the thread that writes that record.
template <int W>
__global__ void __launch_bounds__(kTileThreads)
k_emit_write(GenomeView g, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ stubmask, KParams kp,
             TableView J, RecordTable rt, uint64_t tile_begin, uint64_t tile_end,
             const unsigned long long* __restrict__ tile_rec_prefix, const unsigned long long* __restrict__ tile_stub_prefix,
             uint64_t records_before, uint64_t stubs_before, uint64_t unit_base, uint64_t first_stub_id,
             uint32_t* __restrict__ out, uint64_t out_units) {
    __shared__ unsigned long long warp_tot[8];
    for (uint64_t tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        uint64_t w = tile * kTileThreads + threadIdx.x;
        uint32_t m = 0, stub = 0;
        if (w * 32 < g.npos) { m = mask[w]; stub = stubmask[w]; }
        uint32_t bits = m | stub;
        // block-exclusive scan of (records | stubs << 32)
        unsigned long long mine = ((unsigned long long)__popc(stub) << 32) | (unsigned)__popc(bits);
        unsigned long long incl = mine;
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        __syncthreads();
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        unsigned long long wbase = 0;
        for (int j = 0; j < wid; ++j) wbase += warp_tot[j];
        unsigned long long excl = wbase + incl - mine;
        uint64_t rec_ord = records_before + tile_rec_prefix[tile - tile_begin] + (excl & 0xFFFFFFFFull);
        uint64_t stub_ord = stubs_before + tile_stub_prefix[tile - tile_begin] + (excl >> 32);
        if (!bits) continue;

        // sequence containing the first position to write (binary search, once per thread)
        uint64_t p_first = w * 32 + (__ffs(bits) - 1);
        uint64_t lo = 0, hi = rt.n;  // largest c with start[c] <= p_first
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) >> 1;
            if (rt.start[mid] <= p_first) lo = mid; else hi = mid;
        }
        uint64_t c = lo;
        uint64_t c_start = rt.start[c];
        uint64_t c_next = (c + 1 < rt.n) ? rt.start[c + 1] : ~0ull;
        while (bits) {
            int i = __ffs(bits) - 1;
            bits &= bits - 1;
            uint64_t p = w * 32 + i;
            while (p >= c_next) { ++c; c_start = c_next; c_next = (c + 1 < rt.n) ? rt.start[c + 1] : ~0ull; }
            long long id;
            if ((stub >> i) & 1u) id = (long long)(first_stub_id + stub_ord++);
            else {
                Occ<W> o = occurrence_at<W>(g, p, kp);
                id = lookup_id<W>(g, J, o, kp.k);
            }
            uint64_t unit = rec_ord + c - unit_base;
            if (p == c_start) {  // first record of sequence c: separators first (junctionapi.h:120-123)
                uint32_t ns = rt.sep_before[c];
                for (uint32_t s = 1; s <= ns; ++s) {
                    uint64_t u = unit - s;
                    if (u < out_units) { out[3 * u] = 0xFFFFFFFFu; out[3 * u + 1] = 0xFFFFFFFFu; out[3 * u + 2] = 0x7FFFFFFFu; }
                }
            }
            if (unit < out_units) {
                out[3 * unit] = (uint32_t)(p - c_start);  // u32 position inside the record (h:938)
                out[3 * unit + 1] = (uint32_t)(unsigned long long)id;
                out[3 * unit + 2] = (uint32_t)((unsigned long long)id >> 32);
            }
            ++rec_ord;
        }
    }
}

// ------------------------------------------------------------------------------------------
// GetId (vertexenumerator.h:98-102) for one k-mer given as packed words (host packs the string)
// ------------------------------------------------------------------------------------------
template <int W>
__global__ void k_get_id(GenomeView g, TableView J, KParams kp, Kmer<W> x, long long* out) {
    Occ<W> o;
    o.X = x;
    o.Y = revcomp<W>(x, kp.k);
    o.fwd = kmer_less<W>(o.X, o.Y);
    o.fold = 0;
    o.h = kmer_hash<W>(kmer_select<W>(o.fwd, o.X, o.Y), kp.seed);
    *out = lookup_id<W>(g, J, o, kp.k);
}

}  // namespace tpc
