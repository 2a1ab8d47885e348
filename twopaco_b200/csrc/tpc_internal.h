// tpc_internal.h -- shared between the translation units of libtwopaco_b200.so (not installed).
#pragma once
#include "../../include/twopaco_b200.h"

namespace tpc {
int set_error(const char* fmt, ...);
const char* last_error();
}  // namespace tpc

extern "C" {
// single-GPU pipeline on a session whose genome is set: everything up to the record count
// (-> size of the image), then the write + device->host copy of the image.
int tpc_session_run_to_count(tpc_session* s, uint64_t* image_bytes);
int tpc_session_write_host(tpc_session* s, uint8_t* out_image, uint64_t image_bytes);
// the same, handing the image to `sink` chunk by chunk (two pinned staging buffers; the device->host
// copy of chunk i+1 overlaps the sink's work on chunk i)
typedef int (*tpc_chunk_sink)(void* ctx, const uint8_t* data, uint64_t nbytes);
int tpc_session_write_stream(tpc_session* s, uint64_t image_bytes, tpc_chunk_sink sink, void* ctx);
}
