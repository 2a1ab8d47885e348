// tpc_launch.cuh -- host-callable launchers, one explicit instantiation per k-mer word count W
// (tpc_w1.cu .. tpc_w4.cu) so that the template expansion compiles in parallel.
#pragma once
#include "tpc_device.cuh"

namespace tpc {

struct Counters;
struct RecordTable;
struct BinView;
struct OwnPlanes;
struct MarkList;

struct LaunchCtx {
    cudaStream_t stream;
    int sm_count;
    uint32_t* launches;  // incremented per kernel launch
    int bin_ctas = 0;    // cap on resident CTAs per SM of the binning kernels (0 = what fits): two kernels bound by
    int apply_ctas = 0;  //   different pipes share the SMs when the rounds are pipelined; same for the apply kernels
};

template <int W>
struct Launch {
    // tiles [tile_begin, tile_end)
    static cudaError_t fill(const LaunchCtx&, GenomeView, uint32_t* filter, KParams, uint64_t tile_begin, uint64_t tile_end, Counters*);
    static cudaError_t query(const LaunchCtx&, GenomeView, const uint32_t* filter, KParams, uint64_t tile_begin, uint64_t tile_end,
                             uint32_t* mask, int accumulate, Counters*, uint32_t* hll);
    static cudaError_t valid_mask(const LaunchCtx&, GenomeView, KParams, uint64_t tile_begin, uint64_t tile_end, uint32_t* mask);
    // ownership planes of the tiles [tile_begin, tile_end): local round (part - part_base, < nlocal) of every position
    static cudaError_t own(const LaunchCtx&, GenomeView, KParams, uint32_t part_base, uint32_t nlocal, uint64_t tile_begin,
                           uint64_t tile_end, const OwnPlanes&);
    // records of the owned positions partitioned by filter slice; planes == nullptr: unsharded (every k-mer is owned)
    static cudaError_t bin(const LaunchCtx&, GenomeView, KParams, const BinView&, uint64_t tile_begin, uint64_t tile_end,
                           uint64_t wave_base, const OwnPlanes* planes);
    static cudaError_t insert(const LaunchCtx&, GenomeView, const uint32_t* mask, KParams, uint64_t tile_begin, uint64_t tile_end,
                              TableView T, Counters*, const OwnPlanes* planes);
    static cudaError_t insert_list(const LaunchCtx&, GenomeView, const MarkList&, KParams, TableView T, Counters*);
    static cudaError_t build_index(const LaunchCtx&, GenomeView, const unsigned long long* sorted, uint64_t n, KParams, TableView J);
    static cudaError_t ends(const LaunchCtx&, GenomeView, const RecordTable&, KParams, TableView J, uint32_t* stubmask,
                            uint64_t pos_begin, uint64_t pos_end);
    static cudaError_t emit_count(const LaunchCtx&, GenomeView, uint32_t* mask, const uint32_t* stubmask, KParams, TableView J,
                                  uint64_t tile_begin, uint64_t tile_end, unsigned long long* tile_records,
                                  unsigned long long* tile_stubs);
    static cudaError_t emit_write(const LaunchCtx&, GenomeView, const uint32_t* mask, const uint32_t* stubmask, KParams,
                                  TableView J, const RecordTable&, uint64_t tile_begin, uint64_t tile_end,
                                  const unsigned long long* tile_rec_prefix, const unsigned long long* tile_stub_prefix,
                                  uint64_t records_before, uint64_t stubs_before, uint64_t unit_base, uint64_t first_stub_id,
                                  uint32_t* out, uint64_t out_units);
    static cudaError_t get_id(const LaunchCtx&, GenomeView, TableView J, KParams, const uint64_t* words, long long* d_out);
};

// W-independent kernels (tpc_session.cu)
cudaError_t launch_classify(const LaunchCtx&, TableView T, uint64_t abundance, uint32_t use_abundance,
                            unsigned long long* out, unsigned long long* out_keys, uint64_t out_cap, Counters*);
cudaError_t launch_scan_exclusive(const LaunchCtx&, unsigned long long* data, uint64_t n, unsigned long long* scratch);
uint64_t scan_scratch_items(uint64_t n);

}  // namespace tpc
