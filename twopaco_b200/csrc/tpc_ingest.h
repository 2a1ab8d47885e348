// tpc_ingest.h -- internal: multi-threaded FASTA ingest (tpc_ingest.cpp)
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace tpc {

struct IngestResult {
    uint8_t* ascii = nullptr;      // pinned host buffer, 1 byte per position in the tpc_genome layout
    uint64_t ascii_bytes = 0;      // allocated size (multiple of 64, >= n_positions + 64)
    uint64_t n_positions = 0;
    bool pinned = false;
    std::vector<uint64_t> rec_start, rec_len;
    ~IngestResult();
};

// Parses all files with `threads` host threads.  Returns 0 or sets the error message
// (same texts as the reference: "Can't open file ...", "Found an invalid character ...").
int ingest_fasta(const char* const* paths, size_t n_files, uint32_t threads, IngestResult* out);

}  // namespace tpc
