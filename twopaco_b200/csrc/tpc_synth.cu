// tpc_synth.cu -- (1) K0: ASCII -> 2-bit codes + N mask on the device (the packing the
// reference never does for the stream: it re-parses ASCII once per stage, h:1135-1214);
// (2) on-device generator of the synthetic founder-family genome sets of SURVEY.md 8(d), so
// that the 3.1 Gbp x 7 benchmark inputs never have to be produced on host cores.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../include/twopaco_b200.h"
#include "tpc_internal.h"
#include "tpc_device.cuh"
#include "tpc_launch.cuh"

using namespace tpc;

#define CKS(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return tpc::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                                  cudaGetErrorString(e_));                                        \
    } while (0)

namespace {

// dnachar.cpp:18-33 (MakeUpChar) after upper-casing; everything else is 'N' (h:1174)
__device__ __forceinline__ uint32_t ascii_code(uint32_t b) {
    b &= 0xDFu;  // fold case
    return b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 4u;
}

// K0: thread t packs positions [32t, 32t+32): one code word and half an n-mask word.
__global__ void __launch_bounds__(256)
k_pack_ascii(const uint8_t* __restrict__ ascii, uint64_t npos, uint64_t* __restrict__ codes, uint64_t code_words,
             uint32_t* __restrict__ nmask_halves, uint64_t half_words) {
    uint64_t n_threads = code_words > half_words ? code_words : half_words;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n_threads; t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p0 = t * 32;
        uint64_t code = 0;
        uint32_t nm = 0;
        if (p0 + 32 <= npos) {
            const uint4* src = reinterpret_cast<const uint4*>(ascii + p0);
            uint4 v[2] = {__ldg(src), __ldg(src + 1)};
            const uint32_t* w = reinterpret_cast<const uint32_t*>(v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                uint32_t c = ascii_code((w[j >> 2] >> (8 * (j & 3))) & 0xFFu);
                if (c < 4) code |= (uint64_t)c << (2 * j); else nm |= 1u << j;
            }
        } else {
            for (int j = 0; j < 32; ++j) {
                uint32_t c = (p0 + j < npos) ? ascii_code(ascii[p0 + j]) : 4u;
                if (c < 4) code |= (uint64_t)c << (2 * j); else nm |= 1u << j;
            }
        }
        if (t < code_words) codes[t] = code;
        if (t < half_words) nmask_halves[t] = nm;
    }
}

// ---- synthetic founder family ----------------------------------------------------------------
struct SynthParams {
    uint64_t seed;
    uint32_t genomes, records;   // records per genome
    uint64_t record_len;         // founder record length
    uint32_t p_threshold;        // p * 2^32
    uint64_t chunks_per_record;
};
constexpr int kSynthPerThread = 8;
constexpr int kSynthChunk = 256 * kSynthPerThread;

__device__ __forceinline__ uint32_t founder_base(uint64_t seed, uint32_t c, uint64_t i) {
    return (uint32_t)(fmix64((seed ^ 0xF00DFACE5EEDull) + (((uint64_t)c << 36) | i)) >> 17) & 3u;
}

// bases produced by founder position i of record c in genome g: n in {0,1,2}, b[0..n)
__device__ __forceinline__ int synth_emit(const SynthParams& sp, uint32_t g, uint32_t c, uint64_t i, uint8_t b[2]) {
    uint32_t fb = founder_base(sp.seed, c, i);
    b[0] = (uint8_t)fb;
    if (g == 0) return 1;
    uint64_t r = fmix64(sp.seed + fmix64(((uint64_t)(g * sp.records + c) << 36) | i));
    if ((uint32_t)r >= sp.p_threshold) return 1;
    uint32_t kind = (uint32_t)(r >> 32) & 0xFFFFu;
    uint32_t extra = (uint32_t)(r >> 48);
    if (kind < 52429u) {            // 80 %: SNP to a uniform other base
        b[0] = (uint8_t)((fb + 1 + extra % 3) & 3u);
        return 1;
    }
    if (kind < 58982u) {            // 10 %: 1-bp insertion of a uniform base after this one
        b[1] = (uint8_t)(extra & 3u);
        return 2;
    }
    return 0;                        // 10 %: 1-bp deletion
}

__global__ void __launch_bounds__(256)
k_synth_count(SynthParams sp, unsigned long long* __restrict__ chunk_sum) {
    __shared__ unsigned long long red[8];
    uint64_t chunk = blockIdx.x;
    uint64_t rec = chunk / sp.chunks_per_record, ch = chunk % sp.chunks_per_record;
    uint32_t g = (uint32_t)(rec / sp.records), c = (uint32_t)(rec % sp.records);
    uint64_t i0 = ch * kSynthChunk + (uint64_t)threadIdx.x * kSynthPerThread;
    unsigned long long n = 0;
    uint8_t b[2];
    for (int j = 0; j < kSynthPerThread; ++j)
        if (i0 + j < sp.record_len) n += synth_emit(sp, g, c, i0 + j, b);
    unsigned long long t = block_sum(n, red);
    // the record's trailing separator belongs to its last chunk
    if (threadIdx.x == 0) chunk_sum[chunk] = t + (ch + 1 == sp.chunks_per_record ? 1 : 0);
}

__global__ void __launch_bounds__(256)
k_synth_write(SynthParams sp, const unsigned long long* __restrict__ chunk_start, uint8_t* __restrict__ ascii) {
    __shared__ unsigned long long warp_tot[8];
    uint64_t chunk = blockIdx.x;
    uint64_t rec = chunk / sp.chunks_per_record, ch = chunk % sp.chunks_per_record;
    uint32_t g = (uint32_t)(rec / sp.records), c = (uint32_t)(rec % sp.records);
    uint64_t i0 = ch * kSynthChunk + (uint64_t)threadIdx.x * kSynthPerThread;
    uint8_t out[2 * kSynthPerThread];
    unsigned n = 0;
    for (int j = 0; j < kSynthPerThread; ++j) {
        if (i0 + j < sp.record_len) {
            uint8_t b[2];
            int m = synth_emit(sp, g, c, i0 + j, b);
            for (int e = 0; e < m; ++e) out[n++] = b[e];
        }
    }
    unsigned long long incl = n;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    unsigned long long base = 1 + chunk_start[chunk];  // position 0 is the leading separator
    for (int j = 0; j < wid; ++j) base += warp_tot[j];
    base += incl - n;
    for (unsigned e = 0; e < n; ++e) ascii[base + e] = "ACGT"[out[e]];
    if (ch + 1 == sp.chunks_per_record && threadIdx.x == 255) ascii[base + n] = 'N';  // separator after the record
}

}  // namespace

extern "C" {

int tpc_pack_ascii_device(const uint8_t* dev_ascii, uint64_t n_positions, uint64_t* dev_codes, uint64_t* dev_nmask, void* stream) {
    if (!dev_ascii || !dev_codes || !dev_nmask) return set_error("null argument");
    if (((uintptr_t)dev_ascii) & 15) return set_error("ascii buffer must be 16-byte aligned");
    uint64_t cw = tpc_code_words(n_positions), hw = 2 * tpc_mask_words(n_positions);
    uint64_t n_threads = std::max(cw, hw);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t blocks = std::min<uint64_t>((n_threads + 255) / 256, (uint64_t)sms * 16);
    k_pack_ascii<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dev_ascii, n_positions, dev_codes, cw,
                                                                     reinterpret_cast<uint32_t*>(dev_nmask), hw);
    CKS(cudaGetLastError());
    return 0;
}

int tpc_synth_family_device(uint64_t seed, uint32_t genomes, uint32_t records_per_genome, uint64_t record_len, double p,
                            uint8_t** dev_ascii, uint64_t* n_positions, uint64_t* rec_start, uint64_t* rec_len) {
    if (!dev_ascii || !n_positions || !genomes || !records_per_genome || !record_len) return set_error("bad argument");
    if (p < 0 || p >= 1) return set_error("mutation rate must be in [0, 1)");
    SynthParams sp{};
    sp.seed = seed; sp.genomes = genomes; sp.records = records_per_genome; sp.record_len = record_len;
    sp.p_threshold = (uint32_t)(p * 4294967296.0);
    sp.chunks_per_record = (record_len + kSynthChunk - 1) / kSynthChunk;
    uint64_t nrec = (uint64_t)genomes * records_per_genome;
    uint64_t nchunks = nrec * sp.chunks_per_record;
    if (nchunks >= (1ull << 31)) return set_error("synthetic set too large");
    unsigned long long *d_sum = nullptr, *d_scratch = nullptr;
    CKS(cudaMalloc(&d_sum, (nchunks + 1) * 8));
    CKS(cudaMalloc(&d_scratch, scan_scratch_items(nchunks + 1) * 8));
    CKS(cudaMemset(d_sum + nchunks, 0, 8));
    uint32_t launches = 0;
    LaunchCtx lc{nullptr, 148, &launches};
    k_synth_count<<<(unsigned)nchunks, 256>>>(sp, d_sum);
    CKS(cudaGetLastError());
    CKS(launch_scan_exclusive(lc, d_sum, nchunks + 1, d_scratch));
    // record boundaries: start of the first chunk of every record (+ grand total)
    std::vector<unsigned long long> starts(nrec + 1);
    for (uint64_t r = 0; r <= nrec; ++r)
        CKS(cudaMemcpyAsync(&starts[r], d_sum + r * sp.chunks_per_record, 8, cudaMemcpyDeviceToHost, nullptr));
    CKS(cudaDeviceSynchronize());
    uint64_t npos = 1 + starts[nrec];
    for (uint64_t r = 0; r < nrec; ++r) {
        if (rec_start) rec_start[r] = 1 + starts[r];
        if (rec_len) rec_len[r] = starts[r + 1] - starts[r] - 1;
    }
    uint8_t* d_ascii = nullptr;
    uint64_t alloc = (npos + 63) / 64 * 64 + 64;
    CKS(cudaMalloc(&d_ascii, alloc));
    CKS(cudaMemset(d_ascii, 'N', alloc));
    k_synth_write<<<(unsigned)nchunks, 256>>>(sp, d_sum, d_ascii);
    CKS(cudaGetLastError());
    CKS(cudaDeviceSynchronize());
    cudaFree(d_sum);
    cudaFree(d_scratch);
    *dev_ascii = d_ascii;
    *n_positions = npos;
    return 0;
}

int tpc_device_alloc(uint64_t bytes, void** out) {
    if (!out) return set_error("null argument");
    CKS(cudaMalloc(out, std::max<uint64_t>(bytes, 16)));
    return 0;
}

void tpc_device_free(void* p) {
    if (p) cudaFree(p);
}

int tpc_copy_to_device(void* dev_dst, const void* host_src, uint64_t bytes) {
    CKS(cudaMemcpy(dev_dst, host_src, bytes, cudaMemcpyHostToDevice));
    return 0;
}

int tpc_copy_to_host(void* host_dst, const void* dev_src, uint64_t bytes) {
    CKS(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
