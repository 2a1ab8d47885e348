// tpc_launch_impl.cuh -- included by tpc_w<W>.cu; defines Launch<W>.
#pragma once
#include "tpc_bin.cuh"
#include "tpc_kernels.cuh"
#include "tpc_launch.cuh"
#include <algorithm>
#include <cstdlib>

namespace tpc {

// persistent grid: one wave of resident CTAs (148 SMs x occupancy), capped by the work
template <typename Kern>
static int persistent_grid(Kern kern, int threads, int sm_count, uint64_t work_items) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
    if (per_sm < 1) per_sm = 1;
    uint64_t g = (uint64_t)per_sm * (uint64_t)sm_count;
    if (g > work_items) g = work_items;
    if (g < 1) g = 1;
    return (int)g;
}

// (the kernels of the long k-mers, W > 4, take q at run time: one instantiation instead of eight)
#define TPC_Q_SWITCH(q, ...)                                   \
    if constexpr (W > 4) { constexpr int Q = 0; __VA_ARGS__; } \
    else switch (q) {                                          \
        case 1: { constexpr int Q = 1; __VA_ARGS__; } break;   \
        case 2: { constexpr int Q = 2; __VA_ARGS__; } break;   \
        case 3: { constexpr int Q = 3; __VA_ARGS__; } break;   \
        case 4: { constexpr int Q = 4; __VA_ARGS__; } break;   \
        case 5: { constexpr int Q = 5; __VA_ARGS__; } break;   \
        case 6: { constexpr int Q = 6; __VA_ARGS__; } break;   \
        case 7: { constexpr int Q = 7; __VA_ARGS__; } break;   \
        default: { constexpr int Q = 8; __VA_ARGS__; } break;  \
    }

template <int W>
cudaError_t Launch<W>::fill(const LaunchCtx& c, GenomeView g, uint32_t* filter, KParams kp, uint64_t tile_begin, uint64_t tile_end,
                            Counters* ctr) {
    if (tile_end <= tile_begin) return cudaSuccess;
    const bool no_list = getenv("TPC_DIRECT_LIST") && atoi(getenv("TPC_DIRECT_LIST")) == 0;   // (tests: the inline kernels)
    if (kp.nparts > 1 && !no_list) {   // sparse ownership: compact the owned positions first (k_direct_list)
        TPC_Q_SWITCH(kp.q, {
            int grid = persistent_grid(k_direct_list<W, Q, false>, kTileThreads, c.sm_count, tile_end - tile_begin);
            k_direct_list<W, Q, false><<<grid, kTileThreads, 0, c.stream>>>(g, filter, kp, tile_begin, tile_end, nullptr, ctr, nullptr);
        });
        ++*c.launches;
        return cudaGetLastError();
    }
    TPC_Q_SWITCH(kp.q, {
        int grid = persistent_grid(k_fill<W, Q>, kTileThreads, c.sm_count, tile_end - tile_begin);
        k_fill<W, Q><<<grid, kTileThreads, 0, c.stream>>>(g, filter, kp, tile_begin, tile_end, ctr);
    });
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::query(const LaunchCtx& c, GenomeView g, const uint32_t* filter, KParams kp, uint64_t tile_begin, uint64_t tile_end,
                             uint32_t* mask, int accumulate, Counters* ctr, uint32_t* hll) {
    if (tile_end <= tile_begin) return cudaSuccess;
    const bool no_list = getenv("TPC_DIRECT_LIST") && atoi(getenv("TPC_DIRECT_LIST")) == 0;
    if (kp.nparts > 1 && !no_list) {
        // (marks are OR-ed in: the mask words of the tiles must have been cleared -- find_candidates and the windowed driver do)
        if (!accumulate) cudaMemsetAsync(mask + tile_begin * kTileThreads, 0, (tile_end - tile_begin) * kTileThreads * 4, c.stream);
        TPC_Q_SWITCH(kp.q, {
            int grid = persistent_grid(k_direct_list<W, Q, true>, kTileThreads, c.sm_count, tile_end - tile_begin);
            k_direct_list<W, Q, true><<<grid, kTileThreads, 0, c.stream>>>(g, const_cast<uint32_t*>(filter), kp, tile_begin, tile_end, mask, ctr, hll);
        });
        ++*c.launches;
        return cudaGetLastError();
    }
    TPC_Q_SWITCH(kp.q, {
        int grid = persistent_grid(k_query<W, Q>, kTileThreads, c.sm_count, tile_end - tile_begin);
        k_query<W, Q><<<grid, kTileThreads, 0, c.stream>>>(g, filter, kp, tile_begin, tile_end, mask, accumulate, ctr, hll);
    });
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::valid_mask(const LaunchCtx& c, GenomeView g, KParams kp, uint64_t tile_begin, uint64_t tile_end, uint32_t* mask) {
    if (tile_end <= tile_begin) return cudaSuccess;
    int grid = persistent_grid(k_valid_mask<W>, kTileThreads, c.sm_count, tile_end - tile_begin);
    k_valid_mask<W><<<grid, kTileThreads, 0, c.stream>>>(g, kp, tile_begin * kTileThreads, tile_end * kTileThreads, mask);
    ++*c.launches;
    return cudaGetLastError();
}

template <int W, int R, bool FUSED>
static void launch_bin_list_v(const LaunchCtx& c, GenomeView g, KParams kp, const BinView& bv, uint64_t tile_begin, uint64_t tile_end,
                              uint64_t wave_base, const OwnPlanes& op) {
    constexpr size_t smem = bin_list_smem_bytes<W>(R);
    // the attribute is per device (a process may hold sessions on several GPUs) and setting it is cheap
    cudaFuncSetAttribute(k_bin_list<W, R, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bin_list<W, R, FUSED>, kTileThreads, smem);
    if (c.bin_ctas > 0 && per_sm > c.bin_ctas) per_sm = c.bin_ctas;
    uint64_t grid = std::min<uint64_t>((uint64_t)(per_sm < 1 ? 1 : per_sm) * c.sm_count, tile_end - tile_begin);
    k_bin_list<W, R, FUSED><<<(int)grid, kTileThreads, smem, c.stream>>>(g, kp, bv, tile_begin, tile_end, wave_base, op);
}
// planes.n == 0: no ownership planes exist, the kernel decides ownership itself (single-round GPUs)
template <int W, int R>
static void launch_bin_list(const LaunchCtx& c, GenomeView g, KParams kp, const BinView& bv, uint64_t tile_begin, uint64_t tile_end,
                            uint64_t wave_base, const OwnPlanes& op) {
    if (op.n == 0) launch_bin_list_v<W, R, true>(c, g, kp, bv, tile_begin, tile_end, wave_base, op);
    else launch_bin_list_v<W, R, false>(c, g, kp, bv, tile_begin, tile_end, wave_base, op);
}

template <int W, int P>
static void launch_own(const LaunchCtx& c, GenomeView g, KParams kp, uint32_t part_base, uint32_t nlocal, uint64_t tile_begin,
                       uint64_t tile_end, const OwnPlanes& op) {
    int grid = persistent_grid(k_own<W, P>, kTileThreads, c.sm_count, tile_end - tile_begin);
    k_own<W, P><<<grid, kTileThreads, 0, c.stream>>>(g, kp, part_base, nlocal, tile_begin * kTileThreads, tile_end * kTileThreads, op);
}

template <int W>
cudaError_t Launch<W>::own(const LaunchCtx& c, GenomeView g, KParams kp, uint32_t part_base, uint32_t nlocal, uint64_t tile_begin,
                           uint64_t tile_end, const OwnPlanes& op) {
    if (tile_end <= tile_begin) return cudaSuccess;
    switch (op.n) {
        case 1: launch_own<W, 1>(c, g, kp, part_base, nlocal, tile_begin, tile_end, op); break;
        case 2: launch_own<W, 2>(c, g, kp, part_base, nlocal, tile_begin, tile_end, op); break;
        case 3: launch_own<W, 3>(c, g, kp, part_base, nlocal, tile_begin, tile_end, op); break;
        default: launch_own<W, 4>(c, g, kp, part_base, nlocal, tile_begin, tile_end, op); break;
    }
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::bin(const LaunchCtx& c, GenomeView g, KParams kp, const BinView& bv, uint64_t tile_begin, uint64_t tile_end,
                           uint64_t wave_base, const OwnPlanes* planes) {
    if (tile_end <= tile_begin) return cudaSuccess;
    const uint64_t ntiles = tile_end - tile_begin;
    if (planes) {
        // dense binning of the owned positions; stage size >= 1.25 x the expected owned positions of a
        // tile (more CTAs per SM when it is small)
        uint32_t expect = (uint32_t)(kTilePos / kp.nparts) * 5 / 4;
        static const int force_r = getenv("TPC_BIN_R") ? atoi(getenv("TPC_BIN_R")) : 0;   // tuning aid
        if (force_r) expect = (uint32_t)force_r * kTileThreads;
        // (a 2048-record stage with 4 CTAs per SM beats a 4096-record stage with 2: the barriers and the global
        // reservations of one CTA are hidden by the others)
        if constexpr (W > 4) {   // long k-mers: one stage size (the others only matter for the throughput of the short ones)
            launch_bin_list<W, 8>(c, g, kp, bv, tile_begin, tile_end, wave_base, *planes);
        } else {
            if (expect <= 4 * kTileThreads) launch_bin_list<W, 4>(c, g, kp, bv, tile_begin, tile_end, wave_base, *planes);
            else if (force_r != 16) launch_bin_list<W, 8>(c, g, kp, bv, tile_begin, tile_end, wave_base, *planes);
            else launch_bin_list<W, 16>(c, g, kp, bv, tile_begin, tile_end, wave_base, *planes);
        }
    } else {
        cudaFuncSetAttribute(k_bin<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBinSmemBytes);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bin<W>, kTileThreads, kBinSmemBytes);
        uint64_t grid = std::min<uint64_t>((uint64_t)(per_sm < 1 ? 1 : per_sm) * c.sm_count, ntiles);
        k_bin<W><<<(int)grid, kTileThreads, kBinSmemBytes, c.stream>>>(g, kp, bv, tile_begin, tile_end, wave_base);
    }
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::insert(const LaunchCtx& c, GenomeView g, const uint32_t* mask, KParams kp, uint64_t tile_begin, uint64_t tile_end,
                              TableView T, Counters* ctr, const OwnPlanes* planes) {
    if (tile_end <= tile_begin) return cudaSuccess;
    OwnPlanes op{};
    if (planes) op = *planes;
    int grid = persistent_grid(k_insert<W>, kTileThreads, c.sm_count, tile_end - tile_begin);
    k_insert<W><<<grid, kTileThreads, 0, c.stream>>>(g, mask, kp, tile_begin, tile_end, T, ctr, op);
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::insert_list(const LaunchCtx& c, GenomeView g, const MarkList& ml, KParams kp, TableView T, Counters* ctr) {
    if (ml.regions == 0) return cudaSuccess;
    int grid = persistent_grid(k_insert_list<W>, 256, c.sm_count, ml.regions);
    k_insert_list<W><<<grid, 256, 0, c.stream>>>(g, ml, kp, T, ctr);
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::build_index(const LaunchCtx& c, GenomeView g, const unsigned long long* sorted, uint64_t n, KParams kp,
                                   TableView J) {
    if (n == 0) return cudaSuccess;
    int grid = persistent_grid(k_build_index<W>, 256, c.sm_count, (n + 255) / 256);
    k_build_index<W><<<grid, 256, 0, c.stream>>>(g, sorted, n, kp, J);
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::ends(const LaunchCtx& c, GenomeView g, const RecordTable& rt, KParams kp, TableView J, uint32_t* stubmask,
                            uint64_t pos_begin, uint64_t pos_end) {
    if (rt.n == 0) return cudaSuccess;
    int grid = persistent_grid(k_ends<W>, 256, c.sm_count, (rt.n + 255) / 256);
    k_ends<W><<<grid, 256, 0, c.stream>>>(g, rt, kp, J, stubmask, pos_begin, pos_end);
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::emit_count(const LaunchCtx& c, GenomeView g, uint32_t* mask, const uint32_t* stubmask, KParams kp,
                                  TableView J, uint64_t tile_begin, uint64_t tile_end, unsigned long long* tile_records,
                                  unsigned long long* tile_stubs) {
    if (tile_end <= tile_begin) return cudaSuccess;
    int grid = persistent_grid(k_emit_count<W>, kTileThreads, c.sm_count, tile_end - tile_begin);
    k_emit_count<W><<<grid, kTileThreads, 0, c.stream>>>(g, mask, stubmask, kp, J, tile_begin, tile_end, tile_records, tile_stubs);
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::emit_write(const LaunchCtx& c, GenomeView g, const uint32_t* mask, const uint32_t* stubmask, KParams kp,
                                  TableView J, const RecordTable& rt, uint64_t tile_begin, uint64_t tile_end,
                                  const unsigned long long* tile_rec_prefix, const unsigned long long* tile_stub_prefix,
                                  uint64_t records_before, uint64_t stubs_before, uint64_t unit_base, uint64_t first_stub_id,
                                  uint32_t* out, uint64_t out_units) {
    if (tile_end <= tile_begin) return cudaSuccess;
    int grid = persistent_grid(k_emit_write<W>, kTileThreads, c.sm_count, tile_end - tile_begin);
    k_emit_write<W><<<grid, kTileThreads, 0, c.stream>>>(g, mask, stubmask, kp, J, rt, tile_begin, tile_end, tile_rec_prefix,
                                                        tile_stub_prefix, records_before, stubs_before, unit_base,
                                                        first_stub_id, out, out_units);
    ++*c.launches;
    return cudaGetLastError();
}

template <int W>
cudaError_t Launch<W>::get_id(const LaunchCtx& c, GenomeView g, TableView J, KParams kp, const uint64_t* words, long long* d_out) {
    Kmer<W> x;
    for (int j = 0; j < W; ++j) x.w[j] = words[j];
    k_get_id<W><<<1, 1, 0, c.stream>>>(g, J, kp, x, d_out);
    ++*c.launches;
    return cudaGetLastError();
}

}  // namespace tpc
