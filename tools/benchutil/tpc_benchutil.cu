// tpc_benchutil.cu -- BENCH / TEST INFRASTRUCTURE ONLY: libtpc_benchutil.so, not part of the product ABI
// (include/twopaco_b200.h) and not linked into libtwopaco_b200.so.
//   * on-device generator of the synthetic founder-family genome sets of SURVEY.md 8(d), so that the
//     3.1 Gbp x 7 benchmark inputs never have to be produced on host cores;
//   * roofline probes of SURVEY.md 8(d): (a) uniform random 32-byte sector touches into a 2^f-bit table
//     in HBM (the bound of the direct filter kernels), (b) random sector touches inside ONE filter slice
//     that fits L2 with the record stream read beside it (the bound of the binned apply kernels).
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "../../twopaco_b200/csrc/tpc_device.cuh"
#include "../../twopaco_b200/csrc/tpc_kernels.cuh"

using namespace tpc;

static thread_local std::string g_err;
static int set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CKS(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                             cudaGetErrorString(e_));                                             \
    } while (0)

namespace {

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


// ------------------------------------------------------------------------------------------
// (a) random-access roofline probe in HBM: uniform random sector touches into a 2^f-bit table
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_probe(uint32_t* __restrict__ table, uint32_t sector_bits, uint32_t mode, uint64_t per_thread, unsigned long long* sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t x = fmix64(tid + 0x1234567ull);
    uint32_t acc = 0;
    for (uint64_t i = 0; i < per_thread; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        uint64_t h = fmix64(x);
        uint32_t* sec = table + ((h >> (64 - sector_bits)) << 3);
        if (mode == 0) {
            Sector v = ld_sector_nc(sec);
            acc += v.w[0] ^ v.w[1] ^ v.w[2] ^ v.w[3] ^ v.w[4] ^ v.w[5] ^ v.w[6] ^ v.w[7];
        } else if (mode == 1) {
            atomicOr(sec + (h & 7), 1u << ((h >> 3) & 31));
        } else {
            uint32_t m = 1u << ((h >> 3) & 31);
            uint32_t cur = __ldcg(sec + (h & 7));
            if ((cur & m) != m) atomicOr(sec + (h & 7), m);
            acc += cur;
        }
    }
    if (acc == 0x9e3779b9u) atomicAdd(sink, 1ull);
}

// ------------------------------------------------------------------------------------------
// (b) in-L2 probe: the access pattern of the binned apply kernels and nothing else.  `rec` holds n 8-byte
// records {mask seed, sector-in-slice | occurrence code << 25} (two arrays of n words, as the binned path lays
// them out); a thread takes U consecutive records per iteration with two 128-bit streaming loads per 4
// records, issues all U sector loads (one 256-bit load each), then
//   mode 0: folds the sector (pure random-sector load rate),
//   mode 1: the query test (query_sector),
//   mode 2: the fill (fill_sector: test, atomicOr only where bits are missing).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_probe_records(uint32_t* __restrict__ rec, uint64_t n, uint64_t distinct, uint32_t sib_mask) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        // `distinct` different k-mers, every one with the same (seed, sector, code) at each of its occurrences
        const uint64_t id = fmix64(i * 0x9E3779B97F4A7C15ull + 12345) % distinct;
        const uint64_t h = fmix64(id ^ 0xABCDEF0123456789ull);
        const uint32_t code = (uint32_t)(h >> 40) & 3u | (((uint32_t)(h >> 44) & 3u) << 3);
        rec[i] = mask_seed(h);
        rec[n + i] = ((uint32_t)(h >> 8) & sib_mask) | (code << 25);
    }
}

template <int U, int MODE>
__global__ void __launch_bounds__(256)
k_probe_slice(uint32_t* __restrict__ slice, const uint32_t* __restrict__ rec, uint64_t n, uint32_t sib_mask, unsigned long long* sink) {
    static_assert(U % 4 == 0, "U is a multiple of 4");
    const uint4* sd = reinterpret_cast<const uint4*>(rec);
    const uint4* w1 = reinterpret_cast<const uint4*>(rec + n);
    const uint64_t nvec = n / U;
    uint32_t acc = 0;
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t seed[U], word[U];
#pragma unroll
        for (int j = 0; j < U / 4; ++j) {
            const uint4 a = __ldcs(sd + v * (U / 4) + j), b = __ldcs(w1 + v * (U / 4) + j);
            seed[4 * j] = a.x; seed[4 * j + 1] = a.y; seed[4 * j + 2] = a.z; seed[4 * j + 3] = a.w;
            word[4 * j] = b.x; word[4 * j + 1] = b.y; word[4 * j + 2] = b.z; word[4 * j + 3] = b.w;
        }
        Sector s[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            uint32_t* sec = slice + ((uint64_t)(word[j] & sib_mask) << 3);
            s[j] = MODE == 2 ? ld_sector_cg(sec) : ld_sector_nc(sec);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (MODE == 0) acc += s[j].w[0] ^ s[j].w[1] ^ s[j].w[2] ^ s[j].w[3] ^ s[j].w[4] ^ s[j].w[5] ^ s[j].w[6] ^ s[j].w[7] ^ seed[j];
            else if (MODE == 1) acc += query_sector(s[j], mask_from_seed<5>(seed[j])) ? 1u : 0u;
            else acc += fill_sector(slice + ((uint64_t)(word[j] & sib_mask) << 3), s[j], mask_from_seed<5>(seed[j]), word[j] >> 25);
        }
    }
    if (acc == 0x9e3779b9u) atomicAdd(sink, 1ull);
}

template <int U>
void launch_probe_slice(int mode, int blocks, uint32_t* slice, const uint32_t* rec, uint64_t n, uint32_t sib_mask, unsigned long long* sink) {
    if (mode == 0) k_probe_slice<U, 0><<<blocks, 256>>>(slice, rec, n, sib_mask, sink);
    else if (mode == 1) k_probe_slice<U, 1><<<blocks, 256>>>(slice, rec, n, sib_mask, sink);
    else k_probe_slice<U, 2><<<blocks, 256>>>(slice, rec, n, sib_mask, sink);
}

}  // namespace

extern "C" {

const char* tpcb_last_error(void) { return g_err.c_str(); }

int tpcb_synth_family_device(uint64_t seed, uint32_t genomes, uint32_t records_per_genome, uint64_t record_len, double p,
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
    unsigned long long* d_sum = nullptr;
    void* d_scratch = nullptr;
    CKS(cudaMalloc(&d_sum, (nchunks + 1) * 8));
    CKS(cudaMemset(d_sum + nchunks, 0, 8));
    k_synth_count<<<(unsigned)nchunks, 256>>>(sp, d_sum);
    CKS(cudaGetLastError());
    size_t scratch_bytes = 0;
    CKS(cub::DeviceScan::ExclusiveSum(nullptr, scratch_bytes, d_sum, d_sum, nchunks + 1));
    CKS(cudaMalloc(&d_scratch, scratch_bytes ? scratch_bytes : 16));
    CKS(cub::DeviceScan::ExclusiveSum(d_scratch, scratch_bytes, d_sum, d_sum, nchunks + 1));
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


// (a) mode 0 = 32-byte loads, 1 = 4-byte atomicOr, 2 = load + conditional atomicOr.  -> sector touches per second
int tpcb_random_access_probe(uint32_t filter_bits, uint32_t mode, uint64_t touches, double* touches_per_s) {
    if (filter_bits < 9 || filter_bits > 40 || mode > 2) return set_error("bad probe arguments");
    uint32_t* table = nullptr;
    unsigned long long* sink = nullptr;
    uint64_t bytes = (1ull << filter_bits) / 8;
    CKS(cudaMalloc(&table, bytes));
    CKS(cudaMalloc(&sink, 8));
    CKS(cudaMemset(table, 0, bytes));
    CKS(cudaMemset(sink, 0, 8));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = sms * 8;
    uint64_t threads = (uint64_t)blocks * 256;
    uint64_t per_thread = std::max<uint64_t>(touches / threads, 1);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_probe<<<blocks, 256>>>(table, filter_bits - 8, mode, std::max<uint64_t>(per_thread / 8, 1), sink);  // warm-up
    cudaEventRecord(a);
    k_probe<<<blocks, 256>>>(table, filter_bits - 8, mode, per_thread, sink);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(table); cudaFree(sink);
    if (e != cudaSuccess) return set_error("probe failed: %s", cudaGetErrorString(e));
    if (touches_per_s) *touches_per_s = (double)(per_thread * threads) / (ms * 1e-3);
    return 0;
}

// (b) `slices` consecutive slices of 2^slice_log2 bytes are touched one after the other (one launch each, as
// the apply kernels do), `records_per_slice` records each, every k-mer occurring `dup` times (C3: 7 genomes).
// mode as k_probe_slice; U = records per thread and iteration (4 or 8); ctas_per_sm = resident 256-thread CTAs.
// -> records (= sector touches) per second over all launches, first (cold) launch excluded.
int tpcb_slice_probe(uint32_t slice_log2, uint32_t slices, uint64_t records_per_slice, uint32_t dup, uint32_t mode, uint32_t U,
                     uint32_t ctas_per_sm, double* touches_per_s) {
    if (slice_log2 < 12 || slice_log2 > 30 || mode > 2 || !slices || !records_per_slice || (U != 4 && U != 8))
        return set_error("bad probe arguments");
    const uint64_t n = records_per_slice / 8 * 8;
    const uint32_t sib_mask = (uint32_t)((1ull << (slice_log2 - 5)) - 1);
    uint32_t *table = nullptr, *rec = nullptr;
    unsigned long long* sink = nullptr;
    const uint64_t slice_bytes = 1ull << slice_log2;
    CKS(cudaMalloc(&table, slice_bytes * (slices + 1)));
    CKS(cudaMalloc(&rec, n * 8 * (slices + 1)));
    CKS(cudaMalloc(&sink, 8));
    CKS(cudaMemset(table, 0, slice_bytes * (slices + 1)));
    CKS(cudaMemset(sink, 0, 8));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (uint32_t s = 0; s <= slices; ++s)
        k_probe_records<<<sms * 8, 256>>>(rec + (uint64_t)s * 2 * n, n, std::max<uint64_t>(n / std::max(dup, 1u), 1), sib_mask);
    CKS(cudaDeviceSynchronize());
    const int blocks = sms * (int)std::max(1u, std::min(ctas_per_sm, 8u));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (uint32_t s = 0; s <= slices; ++s) {
        if (s == 1) cudaEventRecord(a);
        uint32_t* sl = table + (uint64_t)s * (slice_bytes / 4);
        const uint32_t* r = rec + (uint64_t)s * 2 * n;
        if (U == 4) launch_probe_slice<4>((int)mode, blocks, sl, r, n, sib_mask, sink);
        else launch_probe_slice<8>((int)mode, blocks, sl, r, n, sib_mask, sink);
    }
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(table); cudaFree(rec); cudaFree(sink);
    if (e != cudaSuccess) return set_error("probe failed: %s", cudaGetErrorString(e));
    if (touches_per_s) *touches_per_s = (double)n * slices / (ms * 1e-3);
    return 0;
}

}  // extern "C"
