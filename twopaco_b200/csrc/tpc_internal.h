// tpc_internal.h -- shared between the translation units of libtwopaco_b200.so (not installed).
#pragma once
#include "../../include/twopaco_b200.h"

#include <cuda_runtime_api.h>

namespace tpc {
int set_error(const char* fmt, ...);
const char* last_error();

// Source of the packed genome of a position-windowed session (tpc_windowed.inl): windows of whole tiles, in ascending
// order within a pass.  fetch() returns VIRTUAL base pointers (word 0 of the whole arrays): valid, once `ready` has
// completed and until release(), for the code words [tile_begin * 256 - 1, tile_end * 256 + 8) and the n-mask words
// [(tile_begin * 256 - 1) / 2, tile_end * 128 + 4) -- what the kernels read around the window's tiles.
struct WindowProvider {
    enum { kShared = 0,    // every shard scans these windows (several GPUs: upload 1/N each, all-gather)
           kPrivate = 1 }; // only this shard reads them (position-sharded emit)
    virtual ~WindowProvider() {}
    virtual void begin_pass(int kind, uint64_t tile_first, uint64_t tile_last) = 0;
    virtual int fetch(uint64_t tile_begin, uint64_t tile_end, const uint64_t** codes_v, const uint64_t** nmask_v, cudaEvent_t* ready) = 0;
    virtual void release(uint64_t tile_begin) = 0;
    // Shards that share the passes (and the collectives inside them) must repeat a round together: true when ANY shard
    // says so.  One shard: its own answer.
    virtual bool any_shard(bool mine) { return mine; }
};
}  // namespace tpc

// receives a window's part of the image (device memory, final) in image order; may enqueue work on `stream`
typedef int (*tpc_window_sink)(void* ctx, const uint8_t* dev_bytes, uint64_t image_offset, uint64_t nbytes, cudaStream_t stream);

extern "C" {
// single-GPU pipeline on a session whose genome is set: everything up to the record count
// (-> size of the image), then the write + device->host copy of the image.
int tpc_session_run_to_count(tpc_session* s, uint64_t* image_bytes);
int tpc_session_write_host(tpc_session* s, uint8_t* out_image, uint64_t image_bytes);
// the same, handing the image to `sink` chunk by chunk (two pinned staging buffers; the device->host
// copy of chunk i+1 overlaps the sink's work on chunk i)
typedef int (*tpc_chunk_sink)(void* ctx, const uint8_t* data, uint64_t nbytes);
int tpc_session_write_stream(tpc_session* s, uint64_t image_bytes, tpc_chunk_sink sink, void* ctx);

// position-windowed sessions (tpc_windowed.inl): the genome never sits in HBM as a whole
int tpc_session_set_genome_windowed(tpc_session* s, uint64_t n_positions, const uint64_t* rec_start, const uint64_t* rec_len,
                                    uint64_t n_records, uint64_t window_tiles, tpc::WindowProvider* provider);
int tpc_session_local_junction_keys(tpc_session* s, const uint64_t** dev_keys);
int tpc_session_set_junctions_keyed(tpc_session* s, const uint64_t* dev_words_all, const uint64_t* dev_keys_all, uint64_t n);
int tpc_session_emit_windowed(tpc_session* s, uint64_t pos_begin, uint64_t pos_end, int write, uint64_t records_before,
                              uint64_t stubs_before, tpc_window_sink sink, void* ctx, uint64_t* n_records, uint64_t* n_stubs);
}
