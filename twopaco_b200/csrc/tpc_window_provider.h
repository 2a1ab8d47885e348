// tpc_window_provider.h -- internal: the packed genome in (pinned) host memory handed to a windowed session window by
// window (tpc_windowed.inl).  Two device buffers per array: while the kernels work on window w, window w+1 is on its way
// (copy engine over PCIe on a side stream).  With several GPUs a window of a SHARED pass (every GPU scans every position:
// filter fill / query) is cut into `world` equal parts, part r uploaded by GPU r over its own PCIe link and the parts
// all-gathered in place over NVLink (`allgather`, NCCL); windows of a PRIVATE pass (position-sharded emit) are uploaded
// whole by the one GPU that reads them.  Host-only code (CUDA runtime API).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>

#include <cuda_runtime_api.h>

#include "tpc_internal.h"

namespace tpc {

class HostWindowProvider : public WindowProvider {
public:
    // allgather(buffer, part_words, stream): in-place all-gather of `world` parts of part_words 64-bit words each (this
    // rank's part already sits at buffer + rank * part_words); unused when world == 1
    using AllGather = std::function<int(uint64_t* buffer, uint64_t part_words, cudaStream_t stream)>;
    using Agree = std::function<bool(bool)>;   // logical OR over the shards (a host-side barrier between their threads)
    void set_agree(Agree a) { agree_ = std::move(a); }
    bool any_shard(bool mine) override { return agree_ ? agree_(mine) : mine; }

    HostWindowProvider(const uint64_t* codes, const uint64_t* nmask, uint64_t n_positions, uint64_t window_tiles, int rank, int world,
                       AllGather allgather)
        : n_positions_(n_positions), window_tiles_(window_tiles), rank_(rank), world_(std::max(world, 1)), allgather_(std::move(allgather)) {
        host_[0] = codes; host_[1] = nmask;
        total_[0] = tpc_code_words(n_positions); total_[1] = tpc_mask_words(n_positions);
    }
    ~HostWindowProvider() override {
        for (int b = 0; b < 2; ++b) {
            for (int a = 0; a < 2; ++a)
                if (buf_[b][a]) cudaFree(buf_[b][a]);
            if (ready_[b]) cudaEventDestroy(ready_[b]);
            if (copied_[b]) cudaEventDestroy(copied_[b]);
        }
        if (h2d_) cudaStreamDestroy(h2d_);
        if (ag_) cudaStreamDestroy(ag_);
    }

    int init() {
        const uint64_t wpt[2] = {256, 128};
        for (int a = 0; a < 2; ++a) cap_[a] = (window_tiles_ * wpt[a] + 32 + world_ - 1) / world_ * world_ + world_;
        for (int b = 0; b < 2; ++b) {
            for (int a = 0; a < 2; ++a)
                if (cudaMalloc((void**)&buf_[b][a], cap_[a] * 8) != cudaSuccess) return set_error("out of device memory for the genome windows");
            if (cudaEventCreateWithFlags(&ready_[b], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&copied_[b], cudaEventDisableTiming) != cudaSuccess)
                return set_error("cudaEventCreate failed");
            held_[b] = kNone;
        }
        if (cudaStreamCreateWithFlags(&h2d_, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ag_, cudaStreamNonBlocking) != cudaSuccess)
            return set_error("cudaStreamCreate failed");
        return 0;
    }

    void begin_pass(int kind, uint64_t tile_first, uint64_t tile_last) override {
        kind_ = kind; pass_first_ = tile_first; pass_last_ = tile_last;
        // (windows still held from the previous pass were uploaded for that pass's kind and range: drop them)
        for (int b = 0; b < 2; ++b)
            if (!in_use_[b]) held_[b] = kNone;
    }

    int fetch(uint64_t tile_begin, uint64_t tile_end, const uint64_t** codes_v, const uint64_t** nmask_v, cudaEvent_t* ready) override {
        int b = held_[0] == tile_begin ? 0 : held_[1] == tile_begin ? 1 : -1;
        if (b < 0) {
            b = !in_use_[0] ? 0 : 1;
            if (in_use_[b]) return set_error("window provider: both buffers are in use");
            if (int rc = issue(b, tile_begin, tile_end)) return rc;
        }
        in_use_[b] = true;
        // the next window of the pass travels while this one is worked on
        const uint64_t next = tile_end;
        if (next < pass_last_ && !in_use_[b ^ 1] && held_[b ^ 1] != next)
            if (int rc = issue(b ^ 1, next, std::min(pass_last_, next + window_tiles_))) return rc;
        *codes_v = buf_[b][0] - first_word_[b][0];
        *nmask_v = buf_[b][1] - first_word_[b][1];
        *ready = ready_[b];
        return 0;
    }

    void release(uint64_t tile_begin) override {
        for (int b = 0; b < 2; ++b)
            if (held_[b] == tile_begin) { in_use_[b] = false; held_[b] = kNone; }
    }

private:
    static constexpr uint64_t kNone = ~0ull;

    int issue(int b, uint64_t t0, uint64_t t1) {
        // code words [t0 * 256 - 2, t1 * 256 + 8), n-mask words from (t0 * 256 - 2) / 2 on: a little before and after the
        // window's tiles, what the kernels read around a tile
        const uint64_t c0 = t0 ? t0 * 256 - 2 : 0, c1 = std::min(total_[0], t1 * 256 + 8);
        const uint64_t m0 = c0 / 2, m1 = std::min(total_[1], t1 * 128 + 6);
        const uint64_t lo[2] = {c0, m0}, hi[2] = {c1, m1};
        const bool shared = kind_ == kShared && world_ > 1 && allgather_;
        for (int a = 0; a < 2; ++a) {
            const uint64_t n = hi[a] - lo[a];
            first_word_[b][a] = lo[a];
            if (n > cap_[a]) return set_error("window provider: window larger than its buffer");
            if (!shared) {
                if (cudaMemcpyAsync(buf_[b][a], host_[a] + lo[a], n * 8, cudaMemcpyHostToDevice, h2d_) != cudaSuccess)
                    return set_error("host to device copy of a genome window failed");
            } else {
                const uint64_t part = (n + world_ - 1) / world_;
                const uint64_t plo = std::min(n, (uint64_t)rank_ * part), phi = std::min(n, plo + part);
                if (phi > plo && cudaMemcpyAsync(buf_[b][a] + plo, host_[a] + lo[a] + plo, (phi - plo) * 8, cudaMemcpyHostToDevice, h2d_) != cudaSuccess)
                    return set_error("host to device copy of a genome window failed");
                if (phi < (uint64_t)(rank_ + 1) * part)   // (padding of the last part)
                    cudaMemsetAsync(buf_[b][a] + phi, 0, ((uint64_t)(rank_ + 1) * part - phi) * 8, h2d_);
            }
        }
        if (cudaEventRecord(copied_[b], h2d_) != cudaSuccess) return set_error("cudaEventRecord failed");
        if (shared) {
            cudaStreamWaitEvent(ag_, copied_[b], 0);
            for (int a = 0; a < 2; ++a) {
                const uint64_t n = hi[a] - lo[a], part = (n + world_ - 1) / world_;
                if (int rc = allgather_(buf_[b][a], part, ag_)) return rc;
            }
            if (cudaEventRecord(ready_[b], ag_) != cudaSuccess) return set_error("cudaEventRecord failed");
        } else {
            if (cudaEventRecord(ready_[b], h2d_) != cudaSuccess) return set_error("cudaEventRecord failed");
        }
        held_[b] = t0;
        return 0;
    }

    const uint64_t* host_[2];
    uint64_t total_[2];
    uint64_t n_positions_, window_tiles_;
    int rank_, world_;
    AllGather allgather_;
    Agree agree_;
    uint64_t* buf_[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    uint64_t cap_[2] = {0, 0};
    uint64_t first_word_[2][2] = {{0, 0}, {0, 0}};
    uint64_t held_[2] = {kNone, kNone};
    bool in_use_[2] = {false, false};
    cudaEvent_t ready_[2] = {nullptr, nullptr}, copied_[2] = {nullptr, nullptr};
    cudaStream_t h2d_ = nullptr, ag_ = nullptr;
    int kind_ = kShared;
    uint64_t pass_first_ = 0, pass_last_ = 0;
};

}  // namespace tpc
