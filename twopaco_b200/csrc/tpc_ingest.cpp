// tpc_ingest.cpp -- multi-threaded FASTA ingest for tpc_build.
//
// The reference parses its input single-threaded, one character per call, and does so again for
// every stage (StreamFastaParser::GetChar, streamfastaparser.cpp:61-93; DistributeTasks,
// vertexenumerator.h:1108-1226): that caps it at ~10-20 Mbp/s whatever the core count.  Here the
// files are parsed ONCE, by all host threads:
//   1. mmap the file; threads collect the positions of every '>' (memchr);
//   2. one cheap sequential walk over those candidates frames the records (a '>' inside a header
//      line belongs to the header; anywhere else it starts a record -- GetChar's rule, cpp:74-77);
//   3. threads count the bases of 4 MiB pieces (whitespace skipped, alphabet validated);
//   4. threads write the normalised bases into ONE pinned buffer in the tpc_genome position layout
//      (1 byte per position, 'N' separators), which is copied to the GPU and packed there by K0
//      (k_pack_ascii).  No per-record strings, no host-side bit packing.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <immintrin.h>

#include "tpc_ingest.h"
#include "tpc_internal.h"

using tpc::set_error;

namespace {

struct Tables {
    uint8_t cls[256];   // 0..3 ACGT, 4 other valid (-> N), 5 whitespace, 6 invalid
    char norm[256];     // normalised letter
    Tables() {
        for (int i = 0; i < 256; ++i) { cls[i] = isspace(i) ? 5 : 6; norm[i] = 'N'; }
        for (const char* p = "ACGTURYKMSWBDHWNXV"; *p; ++p) {   // dnachar.cpp:11
            cls[(unsigned char)*p] = 4;
            cls[(unsigned char)tolower(*p)] = 4;                // GetChar upper-cases (cpp:79-88)
        }
        const char* acgt = "ACGT";
        for (int i = 0; i < 4; ++i) {
            cls[(unsigned char)acgt[i]] = (uint8_t)i; cls[(unsigned char)tolower(acgt[i])] = (uint8_t)i;
            norm[(unsigned char)acgt[i]] = acgt[i]; norm[(unsigned char)tolower(acgt[i])] = acgt[i];
        }
    }
};
const Tables kT;

// ---- byte classification, 32 bytes at a time (AVX2, chosen at run time; the scalar loops remain for tails and
// for hosts without AVX2).  A byte is a sequence letter iff (lo_lut[low nibble] & hi_lut[high nibble]) has bit 0 or 1:
//   bit 0: high nibble 4/6 with low nibble of A B C D G H K M N      bit 1: high nibble 5/7 with low nibble of R S T U V W X Y
//   bit 2: high nibble 0 with 9..D (\t \n \v \f \r)                  bit 3: 0x20 (space)          0: invalid character
// -- exactly the classes of Tables::cls (dnachar.cpp:11 alphabet, either case; isspace() in the C locale).
#define TPC_AVX2 __attribute__((target("avx2,popcnt")))
TPC_AVX2 inline __m256i classify32(__m256i v) {
    const __m256i lo_lut = _mm256_setr_epi8(8, 1, 3, 3, 3, 2, 2, 3, 3, 6, 4, 5, 4, 5, 1, 0, 8, 1, 3, 3, 3, 2, 2, 3, 3, 6, 4, 5, 4, 5, 1, 0);
    const __m256i hi_lut = _mm256_setr_epi8(4, 0, 8, 0, 1, 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 4, 0, 8, 0, 1, 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i nib = _mm256_set1_epi8(0x0F);
    const __m256i lo = _mm256_shuffle_epi8(lo_lut, _mm256_and_si256(v, nib));
    const __m256i hi = _mm256_shuffle_epi8(hi_lut, _mm256_and_si256(_mm256_srli_epi16(v, 4), nib));
    return _mm256_and_si256(lo, hi);
}

// bases and validity of p[0, n): returns the number of whole 32-byte blocks consumed
TPC_AVX2 size_t count_avx2(const unsigned char* p, size_t n, uint64_t* bases, bool* bad) {
    const __m256i zero = _mm256_setzero_si256(), three = _mm256_set1_epi8(3);
    uint64_t cnt = 0;
    uint32_t invalid = 0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i t = classify32(_mm256_loadu_si256((const __m256i*)(p + i)));
        const uint32_t not_letter = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(t, three), zero));
        invalid |= (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(t, zero));
        cnt += (uint64_t)__builtin_popcount(~not_letter);
    }
    *bases += cnt;
    *bad |= invalid != 0;
    return i;
}

// normalised letters of p[0, n) appended at dst (whitespace and anything else skipped, like the scalar loop):
// returns the whole blocks consumed, *dst_io advanced.  Never writes a byte beyond the letters it appends.
TPC_AVX2 size_t emit_avx2(const unsigned char* p, size_t n, uint8_t** dst_io) {
    const __m256i zero = _mm256_setzero_si256(), three = _mm256_set1_epi8(3), upper = _mm256_set1_epi8((char)0xDF);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T'),
                  cN = _mm256_set1_epi8('N');
    uint8_t* dst = *dst_io;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(p + i));
        const __m256i t = classify32(v);
        uint32_t keep = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(t, three), zero));
        const __m256i u = _mm256_and_si256(v, upper);
        const __m256i acgt = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                             _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
        const __m256i out = _mm256_blendv_epi8(cN, u, acgt);
        if (keep == 0xFFFFFFFFu) {
            _mm256_storeu_si256((__m256i*)dst, out);
            dst += 32;
            continue;
        }
        alignas(32) uint8_t tmp[32];
        _mm256_store_si256((__m256i*)tmp, out);
        while (keep) {   // runs of letters between the skipped bytes (normally: one line break per block at most)
            const uint32_t s = (uint32_t)__builtin_ctz(keep);
            const uint32_t rest = keep >> s;
            const uint32_t len = rest == (0xFFFFFFFFu >> s) ? 32 - s : (uint32_t)__builtin_ctz(~rest);
            memcpy(dst, tmp + s, len);
            dst += len;
            keep = s + len >= 32 ? 0u : keep & (0xFFFFFFFFu << (s + len));
        }
    }
    *dst_io = dst;
    return i;
}

const bool kHaveAvx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt") && !getenv("TPC_INGEST_SCALAR");

struct Mapped {
    const unsigned char* p = nullptr;
    size_t n = 0;
    ~Mapped() { if (p && n) munmap((void*)p, n); }
};

struct Piece {      // a byte range of one record's sequence region
    uint32_t file;
    uint64_t rec;   // global record index
    size_t lo, hi;
    uint64_t bases = 0, base_off = 0;
    bool last = false;   // last piece of its record: also covers the separator after the record
};

// bytes of sequence text per work unit (TPC_INGEST_PIECE shrinks it so tests can exercise seams)
size_t piece_bytes() {
    const char* e = getenv("TPC_INGEST_PIECE");
    long v = e ? atol(e) : 0;
    return v > 0 ? (size_t)v : (size_t)(4u << 20);
}

template <typename F>
void parallel_for(size_t n, uint32_t threads, F&& f) {
    std::atomic<size_t> next{0};
    auto run = [&]() { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < threads && t < n; ++t) pool.emplace_back(run);
    run();
    for (auto& t : pool) t.join();
}

}  // namespace

namespace tpc {

struct IngestPlan::Impl {
    std::vector<Mapped> files;
    std::vector<Piece> pieces;
};
IngestPlan::IngestPlan() : impl(new Impl()) {}
IngestPlan::~IngestPlan() {}

int ingest_plan(const char* const* paths, size_t n_files, uint32_t threads, IngestPlan* out) {
    threads = std::max<uint32_t>(1, std::min<uint32_t>(threads, 256));
    std::vector<Mapped>& files = out->impl->files;
    std::vector<Piece>& pieces = out->impl->pieces;
    files.resize(n_files);
    struct Rec { uint32_t file; size_t hdr; };   // header start (for error messages)
    std::vector<Rec> recs;
    const size_t kPiece = piece_bytes();

    for (size_t fi = 0; fi < n_files; ++fi) {
        int fd = open(paths[fi], O_RDONLY);
        if (fd < 0) return set_error("Can't open file %s", paths[fi]);
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); return set_error("Can't open file %s", paths[fi]); }
        Mapped& m = files[fi];
        m.n = (size_t)st.st_size;
        if (m.n) {
            void* p = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { close(fd); m.n = 0; return set_error("Can't map file %s", paths[fi]); }
            madvise(p, m.n, MADV_SEQUENTIAL);
            m.p = (const unsigned char*)p;
        }
        close(fd);
        if (!m.n) continue;
        if (m.p[0] != '>') return set_error("The FASTA header should start with a '>', started with '%c'", m.p[0]);
        // 1. candidates: every '>' of the file, found by all threads
        const size_t nchunks = std::max<size_t>(1, std::min<size_t>(threads * 4, m.n / (1 << 20) + 1));
        std::vector<std::vector<size_t>> found(nchunks);
        parallel_for(nchunks, threads, [&](size_t c) {
            size_t lo = m.n * c / nchunks, hi = m.n * (c + 1) / nchunks;
            const unsigned char* q = m.p + lo;
            while (q < m.p + hi) {
                q = (const unsigned char*)memchr(q, '>', (size_t)(m.p + hi - q));
                if (!q) break;
                found[c].push_back((size_t)(q - m.p));
                ++q;
            }
        });
        // 2. framing: a candidate inside a header line is header text
        size_t header_end = 0;   // first byte after the newline of the current header
        size_t prev_seq_lo = 0;
        bool open_rec = false;
        auto close_record = [&](size_t seq_hi) {
            uint64_t r = recs.size() - 1;
            for (size_t lo = prev_seq_lo;; lo += kPiece) {
                bool last = lo + kPiece >= seq_hi;
                pieces.push_back(Piece{(uint32_t)fi, r, lo, std::min(seq_hi, lo + kPiece), 0, 0, last});
                if (last) break;
            }
        };
        for (auto& v : found) {
            for (size_t pos : v) {
                if (pos < header_end) continue;          // '>' inside a header line
                if (open_rec) close_record(pos);
                recs.push_back(Rec{(uint32_t)fi, pos});
                const unsigned char* nl = (const unsigned char*)memchr(m.p + pos, '\n', m.n - pos);
                header_end = nl ? (size_t)(nl - m.p) + 1 : m.n;
                prev_seq_lo = header_end;
                open_rec = true;
            }
        }
        if (open_rec) close_record(m.n);
    }

    // 3. bases per piece (+ alphabet check)
    std::atomic<long long> bad_piece{-1};
    parallel_for(pieces.size(), threads, [&](size_t i) {
        Piece& pc = pieces[i];
        const unsigned char* p = files[pc.file].p;
        uint64_t n = 0;
        bool bad = false;
        size_t b = pc.lo;
        if (kHaveAvx2) b += count_avx2(p + b, pc.hi - b, &n, &bad);
        for (; b < pc.hi; ++b) {
            uint8_t c = kT.cls[p[b]];
            n += c <= 4;
            bad |= c == 6;
        }
        pc.bases = n;
        if (bad) {
            long long expect = -1;
            long long mine = (long long)i;
            while (!bad_piece.compare_exchange_weak(expect, mine) && (expect < 0 || expect > mine)) {}
        }
    });
    if (bad_piece.load() >= 0) {
        const Piece& pc = pieces[(size_t)bad_piece.load()];
        const unsigned char* p = files[pc.file].p;
        for (size_t b = pc.lo; b < pc.hi; ++b) {
            if (kT.cls[p[b]] == 6) {
                const Rec& r = recs[pc.rec];
                size_t h0 = r.hdr + 1, h1 = h0;
                while (h1 < files[r.file].n && !isspace(files[r.file].p[h1])) ++h1;
                std::string name((const char*)files[r.file].p + h0, h1 - h0);
                return set_error("Found an invalid character '%c' in sequence %s", p[b], name.c_str());
            }
        }
    }
    out->rec_len.assign(recs.size(), 0);
    for (Piece& pc : pieces) { pc.base_off = out->rec_len[pc.rec]; out->rec_len[pc.rec] += pc.bases; }
    out->rec_start.resize(recs.size());
    uint64_t pos = 1;
    for (size_t r = 0; r < recs.size(); ++r) {
        if (out->rec_len[r] >> 32) return set_error("sequence %zu is longer than 2^32 bp", r);
        out->rec_start[r] = pos;
        pos += out->rec_len[r] + 1;
    }
    out->n_positions = pos;
    out->layout_bytes = (pos + 63) / 64 * 64 + 64;
    return 0;
}

// 4. normalised bases into the position layout, span by span; separators and padding are 'N'.
// Consecutive pieces are contiguous in the layout: a piece covers its bases plus, when it is the
// last piece of its record, the separator that follows the record.
int ingest_emit(const IngestPlan& plan, uint32_t threads, uint64_t span_bytes, IngestSink& sink) {
    threads = std::max<uint32_t>(1, std::min<uint32_t>(threads, 256));
    const std::vector<Mapped>& files = plan.impl->files;
    const std::vector<Piece>& pieces = plan.impl->pieces;
    auto dest_lo = [&](const Piece& pc) { return plan.rec_start[pc.rec] + pc.base_off; };
    auto dest_n = [&](const Piece& pc) { return pc.bases + (pc.last ? 1 : 0); };
    const uint64_t max_piece = piece_bytes() + 1;
    span_bytes = std::max<uint64_t>(span_bytes, 2 * max_piece);
    size_t i = 0;
    uint64_t span_lo = 0;   // layout offset of the current span; position 0 is the leading separator
    while (true) {
        size_t j = i;
        uint64_t span_hi = i < pieces.size() ? dest_lo(pieces[i]) : plan.n_positions;
        while (j < pieces.size() && dest_lo(pieces[j]) + dest_n(pieces[j]) - span_lo <= span_bytes) {
            span_hi = dest_lo(pieces[j]) + dest_n(pieces[j]);
            ++j;
        }
        const bool final_span = j == pieces.size();
        if (final_span) span_hi = plan.layout_bytes;        // trailing padding
        if (span_hi - span_lo > span_bytes + 128 && !final_span) return set_error("ingest: span planning failed");
        uint8_t* buf = sink.acquire(span_hi - span_lo);
        if (!buf) return set_error("ingest: no staging buffer");
        if (span_lo == 0) buf[0] = 'N';
        if (final_span) {
            uint64_t tail_from = std::max<uint64_t>(plan.n_positions, span_lo);
            memset(buf + (tail_from - span_lo), 'N', span_hi - tail_from);
        }
        parallel_for(j - i, threads, [&](size_t t) {
            const Piece& pc = pieces[i + t];
            const unsigned char* p = files[pc.file].p;
            uint8_t* dst = buf + (dest_lo(pc) - span_lo);
            size_t b = pc.lo;
            if (kHaveAvx2) b += emit_avx2(p + b, pc.hi - b, &dst);
            for (; b < pc.hi; ++b) {
                unsigned char ch = p[b];
                if (kT.cls[ch] <= 4) *dst++ = (uint8_t)kT.norm[ch];
            }
            if (pc.last) *dst = 'N';
        });
        if (int rc = sink.commit(span_lo, buf, span_hi - span_lo)) return rc;
        if (final_span) break;
        span_lo = span_hi;
        i = j;
    }
    return 0;
}

}  // namespace tpc

extern "C" {

int tpc_ingest_fasta(const char* const* paths, size_t n_files, uint32_t threads, uint8_t** ascii, uint64_t* n_positions,
                     uint64_t** rec_start, uint64_t** rec_len, uint64_t* n_records) {
    if (!paths || !ascii || !n_positions || !rec_start || !rec_len || !n_records) return set_error("null argument");
    tpc::IngestPlan plan;
    if (int rc = tpc::ingest_plan(paths, n_files, threads, &plan)) return rc;
    size_t n = plan.rec_len.size();
    struct WholeBuffer : tpc::IngestSink {
        uint8_t* base = nullptr;
        uint64_t next = 0;
        uint8_t* acquire(uint64_t) override { return base + next; }
        int commit(uint64_t off, uint8_t*, uint64_t nbytes) override { next = off + nbytes; return 0; }
    } sink;
    sink.base = (uint8_t*)malloc(plan.layout_bytes);
    uint64_t* s = (uint64_t*)malloc(std::max<size_t>(n, 1) * 8);
    uint64_t* l = (uint64_t*)malloc(std::max<size_t>(n, 1) * 8);
    if (!sink.base || !s || !l) { free(sink.base); free(s); free(l); return set_error("out of host memory"); }
    // small spans here so that the tests exercise the span logic
    if (int rc = tpc::ingest_emit(plan, threads, 1, sink)) { free(sink.base); free(s); free(l); return rc; }
    if (n) { memcpy(s, plan.rec_start.data(), n * 8); memcpy(l, plan.rec_len.data(), n * 8); }
    *ascii = sink.base; *n_positions = plan.n_positions; *rec_start = s; *rec_len = l; *n_records = n;
    return 0;
}

void tpc_host_free(void* p) { free(p); }

}  // extern "C"
