// tpc_kernels_common.cuh -- kernels that do not depend on the k-mer word count
// (included by tpc_session.cu only).
#pragma once
#include "tpc_bin.cuh"
#include "tpc_kernels.cuh"

namespace tpc {

__device__ __forceinline__ bool meta_is_junction(unsigned long long meta) {
    int indeg = __popc((uint32_t)meta & 0xFu) + ((meta & kMetaInN2) ? 2 : (meta & kMetaInN1) ? 1 : 0);
    int outdeg = __popc((uint32_t)(meta >> 4) & 0xFu) + ((meta & kMetaOutN2) ? 2 : (meta & kMetaOutN1) ? 1 : 0);
    return indeg > 1 || outdeg > 1;
}

// TrueBifurcations (h:1228-1256): junction && Count <= abundance -> append its word
__global__ void __launch_bounds__(256)
k_classify(TableView T, uint64_t abundance, uint32_t use_abundance, unsigned long long* __restrict__ out,
           unsigned long long* __restrict__ out_keys /* inline tables only; may be null */, uint64_t out_cap, Counters* ctr) {
    uint64_t cap = 1ull << T.log2cap;
    const int lane = threadIdx.x & 31;
    // whole warps iterate together (the bound is rounded up to a warp), so every shuffle below is full-mask
    const uint64_t cap_w = (cap + 31) & ~31ull;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap_w; i += (uint64_t)gridDim.x * blockDim.x) {
        ulonglong2 v = make_ulonglong2(0ull, 0ull);
        if (i < cap) v = *reinterpret_cast<const ulonglong2*>(T.slots + i);
        bool j = v.x != 0 && meta_is_junction(v.y);
        bool drop = j && use_abundance && (v.y >> kMetaCountShift) > abundance;
        if (drop) atomicAdd(&ctr->dropped, 1ull);
        bool take = j && !drop;
        // warp-aggregated append: one atomicAdd per warp, the base handed out by a full-mask shuffle
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        if (ballot == 0) continue;
        const int leader = __ffs(ballot) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(&ctr->junctions, (unsigned long long)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (take) {
            unsigned long long at = base + __popc(ballot & ((1u << lane) - 1));
            if (at < out_cap) {
                out[at] = T.inline_keys ? (v.y >> (kInlinePosShift + 1)) : (v.x & kPosMask);
                if (out_keys) out_keys[at] = v.x | (((v.y >> kInlinePosShift) & 1ull) << 63);   // the J slot's first word
            }
        }
    }
}

// junction index J (bifurcationstorage.h:27-66).  The same from the junction keys themselves ({canonical k-mer + 1 | strand of the first occurrence << 63}, in the
// order of the sorted first positions): no genome access -- the windowed runs, whose genome is not resident.
__global__ void __launch_bounds__(256)
k_build_index_keys(const unsigned long long* __restrict__ sorted_keys, uint64_t n, KParams kp, TableView J) {
    const uint64_t capmask = (1ull << J.log2cap) - 1;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long mine = sorted_keys[i];
        Kmer<1> canon;
        canon.w[0] = (mine & ~(1ull << 63)) - 1ull;
        const uint64_t h = kmer_hash<1>(canon, kp.seed);
        for (uint64_t idx = hash_slot(h, J.log2cap);; idx = (idx + 1) & capmask) {
            if (atomicCAS(&J.slots[idx].rep, 0ull, mine) == 0ull) {
                J.slots[idx].meta = i + 1;
                break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// exclusive prefix sum over 64-bit counters (tile counts): 2048 items per CTA
// ------------------------------------------------------------------------------------------
constexpr int kScanItems = 8;
constexpr int kScanBlock = 256 * kScanItems;

__global__ void __launch_bounds__(256)
k_scan_reduce(const unsigned long long* __restrict__ in, uint64_t n, unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long red[8];
    uint64_t base = (uint64_t)blockIdx.x * kScanBlock;
    unsigned long long s = 0;
    for (int j = 0; j < kScanItems; ++j) {
        uint64_t i = base + (uint64_t)j * 256 + threadIdx.x;
        if (i < n) s += in[i];
    }
    unsigned long long t = block_sum(s, red);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
}

// in-place exclusive scan of one CTA-sized chunk, offset by block_prefix[blockIdx.x]
__global__ void __launch_bounds__(256)
k_scan_apply(unsigned long long* __restrict__ data, uint64_t n, const unsigned long long* __restrict__ block_prefix) {
    __shared__ unsigned long long warp_tot[8];
    uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems;
    unsigned long long v[kScanItems];
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) { v[j] = (base + j < n) ? data[base + j] : 0ull; sum += v[j]; }
    unsigned long long incl = sum;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    unsigned long long run = block_prefix ? block_prefix[blockIdx.x] : 0ull;
    for (int j = 0; j < wid; ++j) run += warp_tot[j];
    run += incl - sum;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if (base + j < n) data[base + j] = run;
        run += v[j];
    }
}

// ---- apply: one launch per slice; the slice stays in L2 ------------------------------------------
// A thread takes kApplyU consecutive records per iteration: the record words come in as 128-bit
// streaming loads, then all kApplyU sector loads (one 256-bit load each, random inside the
// L2-resident slice) are in flight together before any of them is consumed.
constexpr int kApplyU = 4;

// Pull the slice into L2 with sequential line prefetches at the start of the launch: a first touch by
// a random 32-byte access costs a whole HBM row activation, a streaming prefetch of the 64 MiB does not.
__device__ __forceinline__ void prefetch_slice(const uint32_t* slice, uint32_t sectors, uint32_t gtid, uint32_t gsize) {
    const char* base = reinterpret_cast<const char*>(slice);
    const size_t bytes = (size_t)sectors * 32;
    for (size_t off = (size_t)gtid * 128; off < bytes; off += (size_t)gsize * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
}

__device__ __forceinline__ uint32_t u4_get(const uint4& v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }

// Record words reach the thread through a private ring of cp.async (LDGSTS) stages in shared memory,
// kApplyDepth iterations ahead: the HBM latency of the record stream is off the critical path, which is
// then just "sector load -> test -> (atomicOr)".  Each thread reads back only what it copied itself, so
// cp.async.wait_group is the only synchronisation.
constexpr int kApplyDepth = 4;
struct ApplyRing {
    uint4 sd[kApplyDepth][256];
    uint4 w1[kApplyDepth][256];
};
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, uint64_t policy) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct ApplyStream {   // per-thread state of the record prefetcher
    const uint4* sd;
    const uint4* w1;
    uint64_t next, nvec;            // 64-bit: a slice may hold more than 2^32 records (small -f, huge inputs)
    uint32_t gsize;
    uint64_t policy;
    int slot;
    __device__ __forceinline__ void issue(ApplyRing& ring, int d) {
        if (next < nvec) {
            cp_async16(&ring.sd[d][threadIdx.x], sd + next, policy);
            cp_async16(&ring.w1[d][threadIdx.x], w1 + next, policy);
        }
        cp_async_commit();
        next += gsize;
    }
    __device__ __forceinline__ void start(ApplyRing& ring, const uint32_t* rec, const uint32_t* rec_b, uint32_t gtid, uint32_t gs, uint64_t nv) {
        sd = reinterpret_cast<const uint4*>(rec); w1 = reinterpret_cast<const uint4*>(rec_b);
        next = gtid; nvec = nv; gsize = gs; slot = 0;
        policy = l2_evict_first_policy();
#pragma unroll
        for (int d = 0; d < kApplyDepth; ++d) issue(ring, d);
    }
    // records of the current iteration; call refill() once their consumers have been issued
    __device__ __forceinline__ void take(ApplyRing& ring, uint4& a, uint4& b) {
        cp_async_wait<kApplyDepth - 1>();
        a = ring.sd[slot][threadIdx.x];
        b = ring.w1[slot][threadIdx.x];
    }
    __device__ __forceinline__ void refill(ApplyRing& ring) {
        issue(ring, slot);
        slot = slot + 1 == kApplyDepth ? 0 : slot + 1;
    }
};

template <int Q>
__global__ void __launch_bounds__(256, 4)
k_apply_fill(uint32_t* __restrict__ slice, const uint32_t* __restrict__ rec, const unsigned long long* __restrict__ count,
             uint64_t cap, uint32_t sib_mask, Counters* ctr) {
    __shared__ unsigned long long red[8];
    __shared__ ApplyRing ring;
    unsigned long long n64 = *count;
    const uint64_t n = n64 > cap ? cap : n64;
    const uint32_t* __restrict__ rec_b = rec + cap;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    uint32_t fresh = 0;
    const uint64_t nvec = n / kApplyU;
    ApplyStream st;
    st.start(ring, rec, rec_b, gtid, gsize, nvec);
    prefetch_slice(slice, sib_mask + 1u, gtid, gsize);
    for (uint64_t v = gtid; v < nvec; v += gsize) {
        uint4 sd, w1;
        st.take(ring, sd, w1);
        uint32_t* sec[kApplyU];
        Sector s[kApplyU];
#pragma unroll
        for (int j = 0; j < kApplyU; ++j) {
            sec[j] = slice + ((uint64_t)(u4_get(w1, j) & sib_mask) << 3);
            s[j] = ld_sector_cg(sec[j]);
        }
        st.refill(ring);
#pragma unroll
        for (int j = 0; j < kApplyU; ++j)
            fresh += fill_sector(sec[j], s[j], mask_from_seed<Q>(u4_get(sd, j)), u4_get(w1, j) >> kBinCodeShift);
    }
    cp_async_wait<0>();
    for (uint64_t i = nvec * kApplyU + gtid; i < n; i += gsize) {
        const uint32_t w1 = __ldcs(rec_b + i);
        fresh += fill_vertex(slice + ((uint64_t)(w1 & sib_mask) << 3), mask_from_seed<Q>(__ldcs(rec + i)), w1 >> kBinCodeShift);
    }
    unsigned long long t = block_sum(fresh, red);
    if (threadIdx.x == 0 && t) atomicAdd(&ctr->filter_new, t);
}

// a candidate: its mask bit, the sketch of distinct candidates, and (when a list is kept) its position appended to the
// CTA's region of the mark list -- the count lives in shared memory during the launch
__device__ __forceinline__ void apply_mark(uint32_t* __restrict__ mask, uint32_t* __restrict__ hll, uint32_t m, uint32_t w1,
                                           uint32_t relpos, uint32_t sib_bits, uint64_t wave_base, uint64_t slice_first_sector,
                                           const MarkList& ml, uint32_t* s_list_count) {
    const uint32_t sib_mask = (1u << sib_bits) - 1u;
    hll_add(hll, m, slice_first_sector | (w1 & sib_mask));
    const uint64_t p = wave_base + (((uint64_t)((w1 & ((1u << kBinCodeShift) - 1u)) >> sib_bits)) << 32) + relpos;
    atomicOr(mask + (p >> 5), 1u << (p & 31));
    if (ml.entries) {
        const uint32_t at = atomicAdd(s_list_count, 1u);
        if (at < ml.region_cap) ml.entries[(uint64_t)blockIdx.x * ml.region_cap + at] = p;
    }
}
__device__ __forceinline__ void mark_list_begin(const MarkList& ml, uint32_t* s_list_count) {
    if (threadIdx.x == 0) *s_list_count = ml.entries && blockIdx.x < ml.regions ? ml.counts[blockIdx.x] : 0u;
    __syncthreads();
}
__device__ __forceinline__ void mark_list_end(const MarkList& ml, uint32_t* s_list_count) {   // call after a CTA-wide barrier
    if (threadIdx.x == 0 && ml.entries && blockIdx.x < ml.regions) ml.counts[blockIdx.x] = *s_list_count;
}

// Candidates are rare (1-2 % of the records) but nearly every warp iteration (128 records) holds one, and a lane that handles its
// mark inline drags the whole warp through the ~60 instructions of apply_mark (position word from HBM, sketch update, mask
// atomicOr, list append) with one or two lanes active: measured, the query kernel ran at 0.79 of the rate of the same loads and
// tests without the marks.  AGG: the warp queues its marks in shared memory -- slots handed out by ballot, the count is a
// warp-uniform register, no atomics -- and handles them 32 at a time with every lane busy.
struct MarkQueueEntry {
    unsigned long long idx;   // record index inside the slice's arrays (its position word is only read when it is a candidate)
    uint32_t w1, m;
};
constexpr int kMarkQueueCap = 64;   // < 32 queued + at most 32 new per ballot

template <int Q, bool AGG>
__global__ void __launch_bounds__(256, 4)
k_apply_query(const uint32_t* __restrict__ slice, const uint32_t* __restrict__ rec, const unsigned long long* __restrict__ count,
              uint64_t cap, uint32_t sib_bits, uint32_t* __restrict__ mask, uint64_t wave_base, Counters* ctr,
              uint32_t* __restrict__ hll, uint64_t slice_first_sector, MarkList ml) {
    __shared__ unsigned long long red[8];
    __shared__ ApplyRing ring;
    __shared__ uint32_t s_list_count;
    __shared__ MarkQueueEntry s_queue[AGG ? 8 : 1][AGG ? kMarkQueueCap : 1];
    mark_list_begin(ml, &s_list_count);
    unsigned long long n64 = *count;
    const uint64_t n = n64 > cap ? cap : n64;
    const uint32_t* __restrict__ rec_b = rec + cap;
    const uint32_t* __restrict__ rec_c = rec + 2 * cap;   // relative positions: read only for the (few) candidates
    const uint32_t sib_mask = (1u << sib_bits) - 1u;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    uint32_t marks = 0;
    const uint64_t nvec = n / kApplyU;
    ApplyStream st;
    st.start(ring, rec, rec_b, gtid, gsize, nvec);
    prefetch_slice(slice, sib_mask + 1u, gtid, gsize);
    if (!AGG) {
        for (uint64_t v = gtid; v < nvec; v += gsize) {
            uint4 sd, w1;
            st.take(ring, sd, w1);
            Sector s[kApplyU];
#pragma unroll
            for (int j = 0; j < kApplyU; ++j) s[j] = ld_sector_nc(slice + ((uint64_t)(u4_get(w1, j) & sib_mask) << 3));
            st.refill(ring);
#pragma unroll
            for (int j = 0; j < kApplyU; ++j) {
                const uint32_t m = mask_from_seed<Q>(u4_get(sd, j));
                if (query_sector(s[j], m)) {
                    apply_mark(mask, hll, m, u4_get(w1, j), __ldcs(rec_c + v * kApplyU + j), sib_bits, wave_base, slice_first_sector, ml, &s_list_count);
                    ++marks;
                }
            }
        }
    } else {
        const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        MarkQueueEntry* q = s_queue[wid];
        uint32_t queued = 0;   // warp-uniform
        // handle the first `cnt` (<= 32) queued marks, one per lane
        auto flush = [&](uint32_t cnt) {
            __syncwarp();
            if (lane < cnt) {
                const MarkQueueEntry e = q[lane];
                apply_mark(mask, hll, e.m, e.w1, __ldcs(rec_c + e.idx), sib_bits, wave_base, slice_first_sector, ml, &s_list_count);
            }
            __syncwarp();
        };
        // whole warps iterate together (lanes past the end carry no record), so that the ballots below are full-mask
        for (uint64_t v0 = (uint64_t)gtid - lane; v0 < nvec; v0 += gsize) {
            const uint64_t v = v0 + lane;
            const bool act = v < nvec;
            uint4 sd = make_uint4(0u, 0u, 0u, 0u), w1 = make_uint4(0u, 0u, 0u, 0u);
            if (act) st.take(ring, sd, w1);
            Sector s[kApplyU];
#pragma unroll
            for (int j = 0; j < kApplyU; ++j) s[j] = ld_sector_nc(slice + ((uint64_t)(u4_get(w1, j) & sib_mask) << 3));
            if (act) st.refill(ring);
#pragma unroll
            for (int j = 0; j < kApplyU; ++j) {
                const uint32_t m = mask_from_seed<Q>(u4_get(sd, j));
                const bool hit = act && query_sector(s[j], m);
                const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                if (bal == 0) continue;
                if (hit) {
                    MarkQueueEntry e;
                    e.idx = v * kApplyU + j; e.w1 = u4_get(w1, j); e.m = m;
                    q[queued + __popc(bal & ((1u << lane) - 1u))] = e;
                    ++marks;
                }
                queued += __popc(bal);
                if (queued >= 32) {
                    flush(32);
                    queued -= 32;
                    if (lane < queued) q[lane] = q[32 + lane];   // the rest (< 32 entries) moves to the front: disjoint halves
                    __syncwarp();
                }
            }
        }
        flush(queued);
    }
    cp_async_wait<0>();
    for (uint64_t i = nvec * kApplyU + gtid; i < n; i += gsize) {
        const uint32_t w1 = __ldcs(rec_b + i);
        const uint32_t m = mask_from_seed<Q>(__ldcs(rec + i));
        if (query_vertex(slice + ((uint64_t)(w1 & sib_mask) << 3), m)) {
            apply_mark(mask, hll, m, w1, __ldcs(rec_c + i), sib_bits, wave_base, slice_first_sector, ml, &s_list_count);
            ++marks;
        }
    }
    unsigned long long t = block_sum(marks, red);
    mark_list_end(ml, &s_list_count);
    if (threadIdx.x == 0 && t) atomicAdd(&ctr->marks, t);
}

// records that did not fit their slice's array (skewed inputs): direct random access
__global__ void __launch_bounds__(256)
k_apply_overflow(uint32_t* __restrict__ filter, const uint32_t* __restrict__ ov, const unsigned long long* __restrict__ ov_count,
                 uint64_t ov_cap, uint32_t sib_bits, uint32_t q, int do_query, uint32_t* __restrict__ mask, uint64_t wave_base,
                 Counters* ctr, uint32_t* __restrict__ hll, MarkList ml) {
    __shared__ unsigned long long red[8];
    __shared__ uint32_t s_list_count;
    mark_list_begin(ml, &s_list_count);
    unsigned long long n = *ov_count;
    if (n > ov_cap) n = ov_cap;
    unsigned long long acc = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        uint4 r = reinterpret_cast<const uint4*>(ov)[i];
        uint32_t* sec = filter + ((((uint64_t)r.w << sib_bits) | (r.y & ((1u << sib_bits) - 1u))) << 3);
        uint32_t m = mask_from_seed_rt(r.x, q);
        if (!do_query) acc += fill_vertex(sec, m, r.y >> kBinCodeShift);
        else if (query_vertex(sec, m)) {
            apply_mark(mask, hll, m, r.y, r.z, sib_bits, wave_base, (uint64_t)r.w << sib_bits, ml, &s_list_count);
            ++acc;
        }
    }
    unsigned long long t = block_sum(acc, red);
    mark_list_end(ml, &s_list_count);
    if (threadIdx.x == 0 && t) atomicAdd(do_query ? &ctr->marks : &ctr->filter_new, t);
}

}  // namespace tpc
