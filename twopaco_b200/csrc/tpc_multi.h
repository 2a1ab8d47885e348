// tpc_multi.h -- internal: entry point of the multi-GPU driver (tpc_multi.cpp) used by tpc_build.
#pragma once
#include "../../include/twopaco_b200.h"

namespace tpc {
// The packed genome sits on GPU 0 of the context (FASTA parsed once, packed by K0 there); it is broadcast chunk by
// chunk to the other GPUs, the shards run, and every GPU pwrite()s its slice of the image to `fd`.  On success
// *session0 is GPU 0's session (kept alive for GetId; it owns nothing of the genome arrays passed in).
int multi_run_from_device0(tpc_multi* m, const tpc_params* params, const uint64_t* dev0_codes, const uint64_t* dev0_nmask,
                           uint64_t n_positions, const uint64_t* rec_start, const uint64_t* rec_len, uint64_t n_records, int fd,
                           uint64_t* image_bytes, tpc_stats* stats, tpc_session** session0);
}  // namespace tpc
