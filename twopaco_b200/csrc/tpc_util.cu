// tpc_util.cu -- K0: ASCII -> 2-bit codes + N mask on the device (the packing the reference never
// does for the stream: it re-parses ASCII once per stage, h:1135-1214), the position-keyed digest of
// a de_bruijn.bin image (parity at sizes no host canonicaliser handles), and the small device-memory
// helpers of the C ABI.
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


// Position-keyed digest of image words: two independent 64-bit sums of mix(global word index, word).  The
// sums of disjoint slices of one image add up to the digest of the whole image, whatever the slicing
// (multi-GPU runs add their slices' digests), and any changed, moved or missing word changes both sums.
__global__ void __launch_bounds__(256)
k_digest(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t word_base, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long red[8];
    unsigned long long a = 0, b = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t gi = word_base + i + 1, w = words[i];
        a += fmix64((gi * 0x9E3779B97F4A7C15ull) ^ w);
        b += fmix64((gi * 0xC2B2AE3D27D4EB4Full) + w * 0x165667B19E3779F9ull);
    }
    unsigned long long ta = block_sum(a, red);
    unsigned long long tb = block_sum(b, red);
    if (threadIdx.x == 0) {
        if (ta) atomicAdd(out, ta);
        if (tb) atomicAdd(out + 1, tb);
    }
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

int tpc_image_digest_device(const uint8_t* dev_image, uint64_t nbytes, uint64_t image_offset, void* stream, uint64_t digest[2]) {
    if (!digest || (nbytes && !dev_image)) return set_error("null argument");
    if ((nbytes | image_offset) & 3 || ((uintptr_t)dev_image & 3)) return set_error("image slices are multiples of 4 bytes");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d = nullptr;
    CKS(cudaMallocAsync(&d, 16, st));
    CKS(cudaMemsetAsync(d, 0, 16, st));
    if (nbytes) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        uint64_t nwords = nbytes / 4;
        uint64_t blocks = std::min<uint64_t>((nwords + 255) / 256, (uint64_t)sms * 8);
        k_digest<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(dev_image), nwords, image_offset / 4, d);
        CKS(cudaGetLastError());
    }
    CKS(cudaMemcpyAsync(digest, d, 16, cudaMemcpyDeviceToHost, st));
    CKS(cudaStreamSynchronize(st));
    CKS(cudaFreeAsync(d, st));
    return 0;
}

int tpc_release_cached_memory(void) {
    int dev = 0;
    CKS(cudaGetDevice(&dev));
    CKS(cudaDeviceSynchronize());
    cudaMemPool_t pool;
    CKS(cudaDeviceGetDefaultMemPool(&pool, dev));
    CKS(cudaMemPoolTrimTo(pool, 0));
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
