// tpc_dump_common.cuh -- shared by the consumers of de_bruijn.bin on the GPU (tpc_dump.cu: seq / group / dot / canonical
// relabelling; tpc_gfa.cu: gfa1 / gfa2 / fasta): unit decoding, decimal text, scans / sorts, record compaction, classes of
// equal keys in first-appearance order.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/twopaco_b200.h"
#include "tpc_internal.h"

using tpc::set_error;

#define CKD(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return tpc::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

namespace {

constexpr unsigned long long kSepId = 0x7FFFFFFFFFFFFFFFull;

struct Unit {   // one 12-byte unit of the image, read as three 32-bit words (the image is only 4-byte aligned)
    uint32_t pos;
    long long id;
};
__device__ __forceinline__ Unit load_unit(const uint32_t* __restrict__ img, uint64_t i) {
    Unit u;
    u.pos = img[3 * i];
    u.id = (long long)((unsigned long long)img[3 * i + 1] | ((unsigned long long)img[3 * i + 2] << 32));
    return u;
}
__device__ __forceinline__ bool is_sep(const Unit& u) { return u.pos == 0xFFFFFFFFu || (unsigned long long)u.id == kSepId; }

__device__ __forceinline__ uint32_t digits_u64(unsigned long long v) {
    uint32_t d = 1;
    while (v >= 10ull) { v /= 10ull; ++d; }
    return d;
}
// writes v in decimal at dst, returns the number of characters
__device__ __forceinline__ uint32_t put_u64(char* dst, unsigned long long v) {
    const uint32_t d = digits_u64(v);
    for (uint32_t i = d; i-- > 0;) { dst[i] = (char)('0' + (v % 10ull)); v /= 10ull; }
    return d;
}
__device__ __forceinline__ uint32_t len_i64(long long v) {
    return v < 0 ? 1 + digits_u64(0ull - (unsigned long long)v) : digits_u64((unsigned long long)v);
}
__device__ __forceinline__ uint32_t put_i64(char* dst, long long v) {
    if (v < 0) { dst[0] = '-'; return 1 + put_u64(dst + 1, 0ull - (unsigned long long)v); }
    return put_u64(dst, (unsigned long long)v);
}

#define GRID_STRIDE(i, n) for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (n); i += (uint64_t)gridDim.x * blockDim.x)

// separator flag (as the scan input) per unit
__global__ void k_sep_flags(const uint32_t* __restrict__ img, uint64_t n, uint32_t* __restrict__ flag) {
    GRID_STRIDE(i, n) flag[i] = is_sep(load_unit(img, i)) ? 1u : 0u;
}
// records (non-separators) compacted: chr / pos / id / unit index, at rec_index = i - (separators before i)
__global__ void k_compact(const uint32_t* __restrict__ img, uint64_t n, const uint32_t* __restrict__ sep_before /* exclusive scan */,
                          uint32_t* __restrict__ chr, uint32_t* __restrict__ pos, long long* __restrict__ id, uint64_t* __restrict__ unit) {
    GRID_STRIDE(i, n) {
        const Unit u = load_unit(img, i);
        if (is_sep(u)) continue;
        const uint64_t r = i - sep_before[i];
        chr[r] = sep_before[i];
        pos[r] = u.pos;
        id[r] = u.id;
        if (unit) unit[r] = i;
    }
}

// ---- seq
__global__ void k_seq_len(const uint32_t* __restrict__ chr, const uint32_t* __restrict__ pos, const long long* __restrict__ id, uint64_t m,
                          unsigned long long* __restrict__ len) {
    GRID_STRIDE(r, m) len[r] = digits_u64(chr[r]) + 1 + digits_u64(pos[r]) + 1 + len_i64(id[r]) + 1;
}
__global__ void k_seq_write(const uint32_t* __restrict__ chr, const uint32_t* __restrict__ pos, const long long* __restrict__ id, uint64_t m,
                            const unsigned long long* __restrict__ off, char* __restrict__ text) {
    GRID_STRIDE(r, m) {
        char* p = text + off[r];
        p += put_u64(p, chr[r]); *p++ = ' ';
        p += put_u64(p, pos[r]); *p++ = ' ';
        p += put_i64(p, id[r]); *p = '\n';
    }
}

// ---- dot (graphdump.cpp:585-606): two edge lines per pair of consecutive records of one sequence
__device__ __forceinline__ uint32_t put_str(char* dst, const char* s) {
    uint32_t n = 0;
    while (s[n]) { dst[n] = s[n]; ++n; }
    return n;
}
__device__ __forceinline__ uint32_t dot_line(char* p, long long from, long long to, const char* color, uint32_t chr, uint32_t pos, bool write) {
    // "\t<from> -> <to>[color=\"<color>\", label=\"chr=<chr> pos=<pos>\"]\n"
    uint32_t n = 0;
    if (!write) {
        uint32_t c = 0;
        while (color[c]) ++c;
        return 1 + len_i64(from) + 4 + len_i64(to) + 8 + c + 14 + digits_u64(chr) + 5 + digits_u64(pos) + 3;
    }
    p[n++] = '\t';
    n += put_i64(p + n, from);
    n += put_str(p + n, " -> ");
    n += put_i64(p + n, to);
    n += put_str(p + n, "[color=\"");
    n += put_str(p + n, color);
    n += put_str(p + n, "\", label=\"chr=");
    n += put_u64(p + n, chr);
    n += put_str(p + n, " pos=");
    n += put_u64(p + n, pos);
    n += put_str(p + n, "\"]\n");
    return n;
}
__global__ void k_dot_len(const uint32_t* __restrict__ chr, const uint32_t* __restrict__ pos, const long long* __restrict__ id, uint64_t m,
                          unsigned long long* __restrict__ len) {
    GRID_STRIDE(r, m) {
        unsigned long long n = 0;
        if (r > 0 && chr[r] == chr[r - 1])
            n = dot_line(nullptr, id[r - 1], id[r], "blue", chr[r - 1], pos[r - 1], false) +
                dot_line(nullptr, -id[r], -id[r - 1], "red", chr[r - 1], pos[r - 1], false);
        len[r] = n;
    }
}
__global__ void k_dot_write(const uint32_t* __restrict__ chr, const uint32_t* __restrict__ pos, const long long* __restrict__ id, uint64_t m,
                            const unsigned long long* __restrict__ off, char* __restrict__ text) {
    GRID_STRIDE(r, m) {
        if (r > 0 && chr[r] == chr[r - 1]) {
            char* p = text + off[r];
            p += dot_line(p, id[r - 1], id[r], "blue", chr[r - 1], pos[r - 1], true);
            dot_line(p, -id[r], -id[r - 1], "red", chr[r - 1], pos[r - 1], true);
        }
    }
}

// ---- group / canon: after the stable sort by key, e = rank in sorted order, idx[e] = record index
__global__ void k_iota(uint32_t* __restrict__ v, uint64_t m) { GRID_STRIDE(i, m) v[i] = (uint32_t)i; }
__global__ void k_abs_keys(const long long* __restrict__ id, uint64_t m, unsigned long long* __restrict__ key) {
    GRID_STRIDE(i, m) key[i] = id[i] < 0 ? 0ull - (unsigned long long)id[i] : (unsigned long long)id[i];
}
// signed ids as radix keys that sort like signed integers (graphdump sorts by GetId(), a signed comparison)
__global__ void k_signed_keys(const long long* __restrict__ id, uint64_t m, unsigned long long* __restrict__ key) {
    GRID_STRIDE(i, m) key[i] = (unsigned long long)id[i] ^ 0x8000000000000000ull;
}
__global__ void k_heads(const unsigned long long* __restrict__ key, uint64_t m, uint32_t* __restrict__ head) {
    GRID_STRIDE(e, m) head[e] = (e == 0 || key[e] != key[e - 1]) ? 1u : 0u;
}
// per class (gid = inclusive scan of head - 1): first sorted rank and first record index
__global__ void k_class_first(const uint32_t* __restrict__ head, const uint32_t* __restrict__ gid_incl, const uint32_t* __restrict__ idx,
                              uint64_t m, uint32_t* __restrict__ first_e, uint32_t* __restrict__ first_idx) {
    GRID_STRIDE(e, m) if (head[e]) {
        const uint32_t g = gid_incl[e] - 1;
        first_e[g] = (uint32_t)e;
        first_idx[g] = idx[e];
    }
}
__global__ void k_group_len(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ chr, const uint32_t* __restrict__ pos, uint64_t m,
                            unsigned long long* __restrict__ len) {
    GRID_STRIDE(e, m) { const uint32_t r = idx[e]; len[e] = digits_u64(chr[r]) + 1 + digits_u64(pos[r]) + 2; }
}
// text length of class order[j] (its members + the newline), in the order the classes are printed
__global__ void k_class_len(const uint32_t* __restrict__ order, const uint32_t* __restrict__ first_e, const unsigned long long* __restrict__ toff,
                            uint64_t groups, uint64_t m, unsigned long long total_text, unsigned long long* __restrict__ clen) {
    GRID_STRIDE(j, groups) {
        const uint32_t g = order[j];
        const unsigned long long lo = toff[first_e[g]], hi = (uint64_t)g + 1 < groups ? toff[first_e[g + 1]] : total_text;
        clen[j] = hi - lo + 1;
    }
}
__global__ void k_class_base(const uint32_t* __restrict__ order, const unsigned long long* __restrict__ cbase_ranked, uint64_t groups,
                             unsigned long long* __restrict__ cbase) {
    GRID_STRIDE(j, groups) cbase[order[j]] = cbase_ranked[j];
}
__global__ void k_group_write(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ head, const uint32_t* __restrict__ gid_incl,
                              const uint32_t* __restrict__ first_e, const unsigned long long* __restrict__ toff,
                              const unsigned long long* __restrict__ cbase, const uint32_t* __restrict__ chr, const uint32_t* __restrict__ pos,
                              uint64_t m, char* __restrict__ text) {
    GRID_STRIDE(e, m) {
        const uint32_t g = gid_incl[e] - 1, r = idx[e];
        char* p = text + cbase[g] + (toff[e] - toff[first_e[g]]);
        p += put_u64(p, chr[r]); *p++ = ' ';
        p += put_u64(p, pos[r]); *p++ = ';'; *p++ = ' ';
        if (e + 1 == m || head[e + 1]) *p = '\n';   // last member of its class
    }
}
// canonical id of every record: rank of its class by first appearance, signed relative to the first occurrence
__global__ void k_canon_ids(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ gid_incl, const uint32_t* __restrict__ first_idx,
                            const uint32_t* __restrict__ class_rank, const long long* __restrict__ id, uint64_t m, long long* __restrict__ out_id) {
    GRID_STRIDE(e, m) {
        const uint32_t g = gid_incl[e] - 1, r = idx[e];
        const long long n = (long long)class_rank[g] + 1;
        const bool same = (id[r] < 0) == (id[first_idx[g]] < 0);
        out_id[r] = same ? n : -n;
    }
}
__global__ void k_class_rank(const uint32_t* __restrict__ order, uint64_t groups, uint32_t* __restrict__ rank) {
    GRID_STRIDE(j, groups) rank[order[j]] = (uint32_t)j;
}
__global__ void k_canon_image(const uint32_t* __restrict__ img, uint64_t n, const uint32_t* __restrict__ sep_before,
                              const long long* __restrict__ canon_id, uint32_t* __restrict__ out) {
    GRID_STRIDE(i, n) {
        const Unit u = load_unit(img, i);
        unsigned long long id = (unsigned long long)u.id;
        uint32_t pos = u.pos;
        if (is_sep(u)) { pos = 0xFFFFFFFFu; id = kSepId; }      // (either field marks a separator: normalise)
        else id = (unsigned long long)canon_id[i - sep_before[i]];
        out[3 * i] = pos; out[3 * i + 1] = (uint32_t)id; out[3 * i + 2] = (uint32_t)(id >> 32);
    }
}

struct Scratch {   // device allocations of one call, released together
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, st);
    }
    template <typename T>
    cudaError_t alloc(T** p, uint64_t count) {
        cudaError_t e = cudaMallocAsync((void**)p, std::max<uint64_t>(count, 1) * sizeof(T), st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    void forget(void* p) { ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), p), ptrs.end()); }
};

int grid_for(uint64_t n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 16));
}

template <typename T>
int exclusive_sum(Scratch& sc, const T* in, T* out, uint64_t n) {
    size_t tmp = 0;
    CKD(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, sc.st));
    void* d = nullptr;
    CKD(sc.alloc((char**)&d, tmp));
    CKD(cub::DeviceScan::ExclusiveSum(d, tmp, in, out, n, sc.st));
    return 0;
}
template <typename T>
int inclusive_sum(Scratch& sc, const T* in, T* out, uint64_t n) {
    size_t tmp = 0;
    CKD(cub::DeviceScan::InclusiveSum(nullptr, tmp, in, out, n, sc.st));
    void* d = nullptr;
    CKD(sc.alloc((char**)&d, tmp));
    CKD(cub::DeviceScan::InclusiveSum(d, tmp, in, out, n, sc.st));
    return 0;
}
template <typename K>
int sort_pairs(Scratch& sc, const K* kin, K* kout, const uint32_t* vin, uint32_t* vout, uint64_t n, int end_bit) {
    size_t tmp = 0;
    CKD(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, n, 0, end_bit, sc.st));
    void* d = nullptr;
    CKD(sc.alloc((char**)&d, tmp));
    CKD(cub::DeviceRadixSort::SortPairs(d, tmp, kin, kout, vin, vout, n, 0, end_bit, sc.st));
    return 0;
}

struct Records {   // the image's records, compacted
    uint64_t n_units = 0, m = 0;
    uint32_t *sep_before = nullptr, *chr = nullptr, *pos = nullptr;
    long long* id = nullptr;
};

int load_records(Scratch& sc, const uint32_t* img, uint64_t n_units, Records* R) {
    R->n_units = n_units;
    uint32_t* flag = nullptr;
    CKD(sc.alloc(&flag, n_units + 1));
    CKD(sc.alloc(&R->sep_before, n_units + 1));
    CKD(cudaMemsetAsync(flag + n_units, 0, 4, sc.st));
    if (n_units) k_sep_flags<<<grid_for(n_units), 256, 0, sc.st>>>(img, n_units, flag);
    if (int rc = exclusive_sum(sc, flag, R->sep_before, n_units + 1)) return rc;
    uint32_t seps = 0;
    CKD(cudaMemcpyAsync(&seps, R->sep_before + n_units, 4, cudaMemcpyDeviceToHost, sc.st));
    CKD(cudaStreamSynchronize(sc.st));
    R->m = n_units - seps;
    CKD(sc.alloc(&R->chr, R->m));
    CKD(sc.alloc(&R->pos, R->m));
    CKD(sc.alloc(&R->id, R->m));
    if (n_units) k_compact<<<grid_for(n_units), 256, 0, sc.st>>>(img, n_units, R->sep_before, R->chr, R->pos, R->id, nullptr);
    CKD(cudaGetLastError());
    return 0;
}

// classes of equal key in first-appearance order: stable sort by key, heads, class ids, classes ranked by first record
struct Classes {
    uint32_t *idx = nullptr, *head = nullptr, *gid_incl = nullptr, *first_e = nullptr, *first_idx = nullptr, *order = nullptr;
    unsigned long long* key_sorted = nullptr;
    uint64_t groups = 0;
};

int build_classes(Scratch& sc, const unsigned long long* key, uint64_t m, Classes* C) {
    if (m >= (1ull << 32)) return set_error("more than 2^32 records in the image");
    uint32_t* iota = nullptr;
    CKD(sc.alloc(&iota, m));
    CKD(sc.alloc(&C->idx, m));
    CKD(sc.alloc(&C->key_sorted, m));
    CKD(sc.alloc(&C->head, m + 1));
    CKD(sc.alloc(&C->gid_incl, m + 1));
    if (m == 0) return 0;
    k_iota<<<grid_for(m), 256, 0, sc.st>>>(iota, m);
    if (int rc = sort_pairs(sc, key, C->key_sorted, iota, C->idx, m, 64)) return rc;   // stable: file order survives inside a class
    k_heads<<<grid_for(m), 256, 0, sc.st>>>(C->key_sorted, m, C->head);
    if (int rc = inclusive_sum(sc, C->head, C->gid_incl, m)) return rc;
    uint32_t groups = 0;
    CKD(cudaMemcpyAsync(&groups, C->gid_incl + (m - 1), 4, cudaMemcpyDeviceToHost, sc.st));
    CKD(cudaStreamSynchronize(sc.st));
    C->groups = groups;
    uint32_t* giota = nullptr;
    uint32_t* fkey_sorted = nullptr;
    CKD(sc.alloc(&C->first_e, groups));
    CKD(sc.alloc(&C->first_idx, groups));
    CKD(sc.alloc(&C->order, groups));
    CKD(sc.alloc(&giota, groups));
    CKD(sc.alloc(&fkey_sorted, groups));
    k_class_first<<<grid_for(m), 256, 0, sc.st>>>(C->head, C->gid_incl, C->idx, m, C->first_e, C->first_idx);
    k_iota<<<grid_for(groups), 256, 0, sc.st>>>(giota, groups);
    if (int rc = sort_pairs(sc, C->first_idx, fkey_sorted, giota, C->order, groups, 32)) return rc;   // classes by first appearance
    CKD(cudaGetLastError());
    return 0;
}

}  // namespace
