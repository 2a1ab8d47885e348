// tpc_ingest.h -- internal: multi-threaded FASTA ingest (tpc_ingest.cpp)
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace tpc {

// Receives the position layout (1 byte per position, 'N' separators and padding) as consecutive
// spans, in order.  acquire() hands out a buffer of at least max_bytes that the parser threads fill;
// commit() is told which layout range it holds (tpc_build: pinned staging -> cudaMemcpyAsync).
struct IngestSink {
    virtual ~IngestSink() {}
    virtual uint8_t* acquire(uint64_t max_bytes) = 0;
    virtual int commit(uint64_t layout_offset, uint8_t* buf, uint64_t nbytes) = 0;
};

struct IngestPlan {
    struct Impl;
    std::unique_ptr<Impl> impl;           // mapped files + pieces
    uint64_t n_positions = 0;
    uint64_t layout_bytes = 0;            // n_positions rounded up to 64 + 64 (what K0 may read)
    std::vector<uint64_t> rec_start, rec_len;
    IngestPlan();
    ~IngestPlan();
};

// Phase 1: map, frame, count and validate with `threads` host threads.  Returns 0 or sets the error
// message (same texts as the reference: "Can't open file ...", "Found an invalid character ...").
int ingest_plan(const char* const* paths, size_t n_files, uint32_t threads, IngestPlan* plan);
// Phase 2: normalise into spans of about span_bytes and hand them to the sink, in layout order.
int ingest_emit(const IngestPlan& plan, uint32_t threads, uint64_t span_bytes, IngestSink& sink);

}  // namespace tpc
