// tpc_gfa.cu -- `graphdump -f gfa1 | gfa2 | fasta` on the GPU (graphdump.cpp:47-113, 175-582).
//
// The reference walks the records of de_bruijn.bin in file order with the input sequences beside them: two consecutive
// records (begin, end) of one sequence are a SEGMENT (a non-branching path of the compacted de Bruijn graph) whose
// id packs the first junction's id, its sign and the first edge character (Segment, graphdump.cpp:47-113); a segment whose
// first edge character is 'N' gets a fresh id from a counter that starts at 2^34.  Per segment it prints, in this order:
// the segment line (only the first time the id is seen -- a 2^35-bit `seen` vector, :400), the occurrence line, the link
// to the previous segment of the sequence; after the last segment of a sequence the path line.
//
// Here every record is one thread: segment ids are independent given an exclusive scan of the 'N'-path flags (the
// counter), "first time seen" is the head of a class after a stable radix sort by |segment id|, and the text position
// of everything follows from prefix sums of exact text lengths computed by the same code that writes the text
// (TextOut: count or write).  Segment bodies (possibly a whole chromosome) are copied by separate warp / CTA kernels.
// Byte-identical to the reference's output, including its quirks: the "s<n>_" prefix counter that never advances
// (:191), non-ACGT edge characters other than 'N' (MakeUpChar = -1 -> segment id -1 / 1), ids of the first sequence
// record being written even when the header line has no token (StreamFastaParser::ReadRecord keeps the previous one).
// Not reproduced (an error here, "The input is corrupted"; undefined behaviour or the same message in the reference,
// :461): images in which a sequence has no record at all (length < k), or that hold more sequences than the FASTA files.
#include <cctype>
#include <map>
#include <string>

#include "tpc_dump_common.cuh"

namespace {

constexpr long long kReservedPath0 = 1ll << 34;        // graphdump.cpp:42-43 (ID_POWER 35)
constexpr unsigned long long kMaxJunctionId = 1ull << 31;   // :44

enum : uint32_t { kGfa1 = 3, kGfa2 = 4, kFasta = 5 };

struct SeqTable {   // the input sequences on the device
    const uint8_t* chars;              // upper-cased characters of all records, back to back
    const unsigned long long* start;   // [n + 1] first character of record c
    const char* id_chars;              // segment names of the records (header token, optionally prefixed), back to back
    const uint32_t* id_start;          // [n + 1]
    uint32_t n;
};

// counts or writes (the same code computes the text length and, after the prefix sums, produces the text)
struct TextOut {
    char* p;
    unsigned long long n = 0;
    __device__ explicit TextOut(char* dst) : p(dst) {}
    __device__ __forceinline__ void ch(char c) { if (p) p[n] = c; ++n; }
    __device__ __forceinline__ void u64(unsigned long long v) { n += p ? put_u64(p + n, v) : digits_u64(v); }
    __device__ __forceinline__ void str(const char* s, uint32_t len) {
        if (p) for (uint32_t i = 0; i < len; ++i) p[n + i] = s[i];
        n += len;
    }
    __device__ __forceinline__ void lit(const char* s) { uint32_t l = 0; while (s[l]) ++l; str(s, l); }
    __device__ __forceinline__ void skip(unsigned long long len) { n += len; }
    // Gfa2Position (graphdump.cpp:268-281)
    __device__ __forceinline__ void gfa2_pos(unsigned long long pos, unsigned long long length) { u64(pos); if (pos == length) ch('$'); }
};

__device__ __forceinline__ unsigned long long abs_ll(long long v) { return v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v; }
__device__ __forceinline__ char sign_of(long long v) { return v >= 0 ? '+' : '-'; }   // Sign (graphdump.cpp:170-173)
__device__ __forceinline__ long long make_up_char(uint8_t c) {   // DnaChar::MakeUpChar (dnachar.cpp:18-33): size_t(-1) otherwise
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}
__device__ __forceinline__ uint8_t reverse_char(uint8_t c) {     // DnaChar::ReverseChar (dnachar.cpp:52-60, 82-85)
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'N';
}
__device__ __forceinline__ bool is_segment(const uint32_t* __restrict__ chr, uint64_t r) { return r > 0 && chr[r] == chr[r - 1]; }

// ---- validation: what the reference calls a corrupted input (graphdump.cpp:459-462) or cannot represent (:56-59)
__global__ void k_gfa_check(const uint32_t* __restrict__ chr, const uint32_t* __restrict__ pos, const long long* __restrict__ id, uint64_t m,
                            SeqTable sq, uint32_t k, uint32_t* __restrict__ bad) {
    GRID_STRIDE(r, m) {
        uint32_t b = 0;
        const uint32_t c = chr[r];
        if (r == 0 ? c != 0 : (c < chr[r - 1] || c - chr[r - 1] > 1)) b |= 1u;
        if (c >= sq.n || (unsigned long long)pos[r] + k > sq.start[min(c, sq.n - 1) + 1] - sq.start[min(c, sq.n - 1)]) b |= 1u;
        if (r > 0 && c == chr[r - 1] && pos[r] <= pos[r - 1]) b |= 1u;
        const bool in_segment = is_segment(chr, r) || (r + 1 < m && chr[r + 1] == c);
        if (in_segment && abs_ll(id[r]) >= kMaxJunctionId) b |= 2u;
        if (b) atomicOr(bad, b);
    }
}

// ---- segment ids (Segment::Segment, graphdump.cpp:50-97).  Record r > 0 of the same sequence as record r - 1 closes the
// segment (begin = r - 1, end = r).  sid[r] = the id, or 0 with uniq[r] = 1 when it comes from the counter.
__global__ void k_gfa_segment_ids(const uint32_t* __restrict__ chr, const uint32_t* __restrict__ pos, const long long* __restrict__ id, uint64_t m,
                                  SeqTable sq, uint32_t k, long long* __restrict__ sid, uint32_t* __restrict__ uniq) {
    GRID_STRIDE(r, m) {
        long long s = 0;
        uint32_t u = 0;
        if (is_segment(chr, r)) {
            const uint8_t* seq = sq.chars + sq.start[chr[r]];
            const uint8_t pos_edge = seq[(unsigned long long)pos[r - 1] + k];
            const uint8_t neg_edge = reverse_char(seq[pos[r] - 1]);
            const long long bid = id[r - 1], eid = id[r];
            const unsigned long long ab = abs_ll(bid), ae = abs_ll(eid);
            const bool fwd = ab < ae || (ab == ae && ab > 0);
            const uint8_t edge = fwd ? pos_edge : neg_edge;
            const long long b_id = fwd ? bid : -eid;
            if (edge == 'N') u = 1;
            else {
                s = make_up_char(edge);
                if (b_id < 0) { s |= 1 << 2; s |= (long long)(abs_ll(b_id) << 3); }
                else s |= b_id << 3;
                if (bid != b_id) s = -s;
            }
        }
        sid[r] = s;
        uniq[r] = u;
    }
}
__global__ void k_gfa_reserved_ids(const uint32_t* __restrict__ uniq, const uint32_t* __restrict__ uniq_before, uint64_t m, long long* __restrict__ sid) {
    GRID_STRIDE(r, m) if (uniq[r]) sid[r] = kReservedPath0 + (long long)uniq_before[r];
}
__global__ void k_gfa_keys(const uint32_t* __restrict__ chr, const long long* __restrict__ sid, uint64_t m, unsigned long long* __restrict__ key) {
    GRID_STRIDE(r, m) key[r] = is_segment(chr, r) ? abs_ll(sid[r]) : ~0ull;
}
// first[r] = 1: the segment closed by record r is the first occurrence of its |id| in file order (-> its segment line)
__global__ void k_gfa_first(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ head, const unsigned long long* __restrict__ key_sorted,
                            uint64_t m, uint8_t* __restrict__ first) {
    GRID_STRIDE(e, m) first[idx[e]] = (head[e] && key_sorted[e] != ~0ull) ? 1 : 0;
}
__global__ void k_gfa_chr_bounds(const uint32_t* __restrict__ chr, uint64_t m, uint32_t* __restrict__ chr_first, uint32_t* __restrict__ chr_last) {
    GRID_STRIDE(r, m) {
        if (r == 0 || chr[r] != chr[r - 1]) chr_first[chr[r]] = (uint32_t)r;
        if (r + 1 == m || chr[r + 1] != chr[r]) chr_last[chr[r]] = (uint32_t)r;
    }
}

struct SegCtx {
    const uint32_t* __restrict__ chr;
    const uint32_t* __restrict__ pos;
    const long long* __restrict__ sid;
    const uint8_t* __restrict__ first;
    SeqTable sq;
    uint32_t k, format;
};

// length of the segment body as printed (FASTA: a line break after every 80 characters and after a last partial line, :532-548)
__device__ __forceinline__ unsigned long long body_text_len(uint32_t format, unsigned long long size) {
    return format == kFasta ? size + (size + 79) / 80 : size;
}
// the part of the segment line before the body
__device__ __forceinline__ void segment_line_head(TextOut& t, uint32_t format, unsigned long long a, unsigned long long size) {
    if (format == kFasta) { t.ch('>'); t.u64(a); t.ch('\n'); return; }
    t.lit("S\t"); t.u64(a); t.ch('\t');
    if (format == kGfa2) { t.u64(size); t.ch('\t'); }
}
// everything record r prints except the body characters and the path line (Gfa1Generator :209-260, Gfa2Generator :289-374)
__device__ __forceinline__ void segment_lines(TextOut& t, const SegCtx& c, uint64_t r) {
    if (!is_segment(c.chr, r)) return;
    const long long s = c.sid[r];
    const unsigned long long a = abs_ll(s), size = (unsigned long long)c.pos[r] + c.k - c.pos[r - 1];
    if (c.first[r]) {
        segment_line_head(t, c.format, a, size);
        t.skip(body_text_len(c.format, size));
        if (c.format != kFasta) t.ch('\n');
    }
    if (c.format == kFasta) return;
    const uint32_t q = c.chr[r];
    const char* name = c.sq.id_chars + c.sq.id_start[q];
    const uint32_t name_len = c.sq.id_start[q + 1] - c.sq.id_start[q];
    const unsigned long long chr_len = c.sq.start[q + 1] - c.sq.start[q];
    if (c.format == kGfa1) {
        t.lit("C\t"); t.u64(a); t.ch('\t'); t.ch(sign_of(s)); t.ch('\t'); t.str(name, name_len); t.lit("\t+\t"); t.u64(c.pos[r]); t.ch('\n');
    } else {
        t.lit("F\t"); t.u64(a); t.ch('\t'); t.str(name, name_len); t.ch(sign_of(s)); t.lit("\t0\t"); t.u64(size); t.lit("$\t");
        t.gfa2_pos(c.pos[r - 1], chr_len); t.ch('\t'); t.gfa2_pos((unsigned long long)c.pos[r] + c.k, chr_len); t.ch('\t'); t.u64(c.k); t.lit("M\n");
    }
    if (r >= 2 && c.chr[r - 2] == q) {   // link to the previous segment of the sequence
        const long long ps = c.sid[r - 1];
        const unsigned long long pa = abs_ll(ps), psize = (unsigned long long)c.pos[r - 1] + c.k - c.pos[r - 2];
        if (c.format == kGfa1) {
            t.lit("L\t"); t.u64(pa); t.ch('\t'); t.ch(sign_of(ps)); t.ch('\t'); t.u64(a); t.ch('\t'); t.ch(sign_of(s)); t.ch('\t'); t.u64(c.k); t.lit("M\n");
        } else {
            const unsigned long long p0 = ps > 0 ? psize - c.k : 0, p1 = ps > 0 ? psize : c.k;
            const unsigned long long s0 = s > 0 ? 0 : size - c.k, s1 = s > 0 ? c.k : size;
            t.lit("E\t"); t.u64(pa); t.ch(sign_of(ps)); t.ch('\t'); t.u64(a); t.ch(sign_of(s)); t.ch('\t');
            t.gfa2_pos(p0, psize); t.ch('\t'); t.gfa2_pos(p1, psize); t.ch('\t'); t.gfa2_pos(s0, size); t.ch('\t'); t.gfa2_pos(s1, size);
            t.ch('\t'); t.u64(c.k); t.lit("M\n");
        }
    }
}
// the segment's entry in the path line of its sequence: "<id><sign>" + the separator unless it is the last one
__device__ __forceinline__ void path_piece(TextOut& t, const SegCtx& c, uint64_t r, bool last) {
    t.u64(abs_ll(c.sid[r])); t.ch(sign_of(c.sid[r]));
    if (!last) t.ch(c.format == kGfa1 ? ',' : ' ');
}
__device__ __forceinline__ void path_head(TextOut& t, const SegCtx& c, uint32_t q) {
    const char* name = c.sq.id_chars + c.sq.id_start[q];
    const uint32_t name_len = c.sq.id_start[q + 1] - c.sq.id_start[q];
    if (c.format == kGfa1) { t.lit("P\t"); t.str(name, name_len); t.ch('\t'); }
    else { t.lit("O\t"); t.str(name, name_len); t.lit("p\t"); }
}
__device__ __forceinline__ void path_tail(TextOut& t, const SegCtx& c) {
    if (c.format == kGfa1) t.lit("\t*\n"); else t.ch('\n');
}

__global__ void k_gfa_len(SegCtx c, uint64_t m, unsigned long long* __restrict__ len_a, unsigned long long* __restrict__ piece) {
    GRID_STRIDE(r, m) {
        TextOut t(nullptr);
        segment_lines(t, c, r);
        len_a[r] = t.n;
        unsigned long long pl = 0;
        if (c.format != kFasta && is_segment(c.chr, r)) {
            TextOut u(nullptr);
            path_piece(u, c, r, r + 1 == m || c.chr[r + 1] != c.chr[r]);
            pl = u.n;
        }
        piece[r] = pl;
    }
}
// total text of record r: its lines + (last record of a sequence that has segments) the path line
__global__ void k_gfa_len_total(SegCtx c, uint64_t m, const unsigned long long* __restrict__ len_a, const unsigned long long* __restrict__ piece_off,
                                const uint32_t* __restrict__ chr_first, unsigned long long* __restrict__ len) {
    GRID_STRIDE(r, m) {
        unsigned long long n = len_a[r];
        const uint32_t q = c.chr[r];
        if (c.format != kFasta && (r + 1 == m || c.chr[r + 1] != q) && chr_first[q] != (uint32_t)r) {
            TextOut t(nullptr);
            path_head(t, c, q);
            path_tail(t, c);
            n += t.n + (piece_off[r + 1] - piece_off[chr_first[q]]);
        }
        len[r] = n;
    }
}
__global__ void k_gfa_write(SegCtx c, uint64_t m, const unsigned long long* __restrict__ off, const unsigned long long* __restrict__ len_a,
                            const unsigned long long* __restrict__ piece_off, const uint32_t* __restrict__ chr_first,
                            const uint32_t* __restrict__ chr_last, char* __restrict__ text) {
    GRID_STRIDE(r, m) {
        TextOut t(text + off[r]);
        segment_lines(t, c, r);
        if (c.format == kFasta || !is_segment(c.chr, r)) continue;
        const uint32_t q = c.chr[r], f = chr_first[q], l = chr_last[q];
        TextOut hc(nullptr);
        path_head(hc, c, q);
        char* path = text + off[l] + len_a[l];                        // the path line follows the lines of the last record
        TextOut pp(path + hc.n + (piece_off[r] - piece_off[f]));
        path_piece(pp, c, r, r == l);
        if (r == l) {
            TextOut h(path);
            path_head(h, c, q);
            TextOut tl(path + hc.n + (piece_off[r + 1] - piece_off[f]));
            path_tail(tl, c);
        }
    }
}

// ---- segment bodies (graphdump.cpp:413-427, 532-548): characters [begin, end + k) of the sequence, or their reverse
// complement when the segment id is negative; FASTA breaks the lines at 80 characters
__device__ __forceinline__ void body_copy(const SegCtx& c, uint64_t r, char* __restrict__ dst, uint32_t lane, uint32_t lanes) {
    const unsigned long long size = (unsigned long long)c.pos[r] + c.k - c.pos[r - 1];
    const uint8_t* __restrict__ src = c.sq.chars + c.sq.start[c.chr[r]] + c.pos[r - 1];
    const bool fwd = c.sid[r] > 0, fasta = c.format == kFasta;
    for (unsigned long long j = lane; j < size; j += lanes) {
        const uint8_t ch = fwd ? src[j] : reverse_char(src[size - 1 - j]);
        dst[fasta ? j + j / 80 : j] = (char)ch;
        if (fasta && (j % 80 == 79 || j + 1 == size)) dst[j + j / 80 + 1] = '\n';
    }
}
constexpr unsigned long long kLongBody = 1ull << 15;
__device__ __forceinline__ char* body_dst(const SegCtx& c, uint64_t r, const unsigned long long* __restrict__ off, char* __restrict__ text,
                                          unsigned long long* size) {
    *size = (unsigned long long)c.pos[r] + c.k - c.pos[r - 1];
    TextOut h(nullptr);
    segment_line_head(h, c.format, abs_ll(c.sid[r]), *size);
    return text + off[r] + h.n;
}
__global__ void __launch_bounds__(256)
k_gfa_bodies_warp(SegCtx c, uint64_t m, const unsigned long long* __restrict__ off, char* __restrict__ text) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r0 = warp * 32; r0 < m; r0 += warps * 32) {
        // 32 records per warp and step: the first-occurrence flags by one coalesced load, then one body after the other
        const uint64_t r = r0 + lane;
        uint32_t todo = __ballot_sync(0xffffffffu, r < m && c.first[r] != 0);
        while (todo) {
            const uint64_t rr = r0 + (__ffs(todo) - 1);
            todo &= todo - 1;
            unsigned long long size;
            char* dst = body_dst(c, rr, off, text, &size);
            if (size < kLongBody) body_copy(c, rr, dst, lane, 32);
        }
    }
}
__global__ void __launch_bounds__(256)
k_gfa_bodies_cta(SegCtx c, const uint32_t* __restrict__ long_list, uint32_t n_long, const unsigned long long* __restrict__ off, char* __restrict__ text) {
    // long bodies (up to a whole chromosome): all CTAs share each of them
    for (uint32_t i = 0; i < n_long; ++i) {
        const uint64_t r = long_list[i];
        unsigned long long size;
        char* dst = body_dst(c, r, off, text, &size);
        body_copy(c, r, dst, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    }
}
__global__ void k_gfa_long_list(SegCtx c, uint64_t m, uint32_t* __restrict__ list, uint32_t* __restrict__ count, uint32_t cap) {
    GRID_STRIDE(r, m) {
        if (!c.first[r]) continue;
        if ((unsigned long long)c.pos[r] + c.k - c.pos[r - 1] >= kLongBody) {
            const uint32_t at = atomicAdd(count, 1u);
            if (at < cap) list[at] = (uint32_t)r;
        }
    }
}

// ---- host side: the input sequences as the reference reads them (ChrReader streamfastaparser.h:140-182,
// ReadInputSequences graphdump.cpp:175-203): upper-cased characters, whitespace dropped, header = first token
struct HostSequences {
    std::vector<uint8_t> chars;
    std::vector<unsigned long long> start{0};
    std::vector<std::string> header, file;
};

int read_sequences(const char* path, HostSequences* hs, std::string* current_header) {
    FILE* f = fopen(path, "rb");
    if (!f) return set_error("Can't open file %s", path);
    static const char* kValid = "ACGTURYKMSWBDHWNXV";   // dnachar.cpp:9-11
    bool valid[256] = {};
    for (const char* p = kValid; *p; ++p) valid[(unsigned char)*p] = true;
    std::vector<char> buf(1 << 20);
    enum { kStart, kHeader, kSeq } state = kStart;
    std::string line;
    bool open_record = false;
    int rc = 0;
    auto close_record = [&]() {
        if (!open_record) return;
        hs->start.push_back(hs->chars.size());
        hs->header.push_back(*current_header);
        hs->file.push_back(path);
        open_record = false;
    };
    size_t got;
    while (rc == 0 && (got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < got && rc == 0; ++i) {
            const unsigned char ch = (unsigned char)buf[i];
            switch (state) {
                case kStart:
                    if (ch != '>') { rc = set_error("The FASTA header should start with a '>', started with '%c'", ch); break; }
                    state = kHeader; line.clear(); open_record = true;
                    break;
                case kHeader:
                    if (ch == '\n') {
                        // `ss >> currentHeader_` (streamfastaparser.cpp:45): a line without a token leaves the previous header
                        size_t a = 0;
                        while (a < line.size() && isspace((unsigned char)line[a])) ++a;
                        size_t b = a;
                        while (b < line.size() && !isspace((unsigned char)line[b])) ++b;
                        if (b > a) *current_header = line.substr(a, b - a);
                        state = kSeq;
                    } else line.push_back((char)ch);
                    break;
                case kSeq:
                    if (isspace(ch)) break;
                    if (ch == '>') { close_record(); state = kHeader; line.clear(); open_record = true; break; }
                    if (!valid[toupper(ch)]) { rc = set_error("Found an invalid character '%c' in sequence %s", ch, current_header->c_str()); break; }
                    hs->chars.push_back((uint8_t)toupper(ch));
                    break;
            }
        }
    }
    fclose(f);
    if (rc == 0) close_record();
    return rc;
}

}  // namespace

extern "C" {

int tpc_graphdump_gfa_device(const uint8_t* dev_image, uint64_t image_bytes, uint32_t format, uint32_t k, const uint8_t* dev_seq_chars,
                             const uint64_t* seq_start, const char* const* seq_name, uint64_t n_seq, void* stream, uint8_t** dev_text,
                             uint64_t* text_bytes) {
    if (!dev_text || !text_bytes || (image_bytes && !dev_image) || (n_seq && (!seq_start || !seq_name))) return set_error("null argument");
    if (format < kGfa1 || format > kFasta) return set_error("format must be 3 (gfa1), 4 (gfa2) or 5 (fasta)");
    if ((uintptr_t)dev_image & 3) return set_error("the image must be 4-byte aligned");
    if (n_seq >= (1ull << 32)) return set_error("too many sequences");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(st);
    const uint32_t* img = reinterpret_cast<const uint32_t*>(dev_image);
    const uint64_t n_units = image_bytes / 12;
    Records R;
    if (int rc = load_records(sc, img, n_units, &R)) return rc;
    const uint64_t m = R.m;
    if (m >= (1ull << 32)) return set_error("more than 2^32 records in the image");
    *dev_text = nullptr;
    *text_bytes = 0;
    if (m == 0) return 0;
    if (n_seq == 0) return set_error("The input is corrupted");

    // sequence table -> device
    std::vector<uint32_t> id_start(n_seq + 1, 0);
    std::string id_chars;
    for (uint64_t c = 0; c < n_seq; ++c) { id_chars += seq_name[c]; id_start[c + 1] = (uint32_t)id_chars.size(); }
    SeqTable sq{};
    unsigned long long* d_start = nullptr;
    char* d_id_chars = nullptr;
    uint32_t* d_id_start = nullptr;
    CKD(sc.alloc(&d_start, n_seq + 1));
    CKD(sc.alloc(&d_id_chars, id_chars.size()));
    CKD(sc.alloc(&d_id_start, n_seq + 1));
    CKD(cudaMemcpyAsync(d_start, seq_start, (n_seq + 1) * 8, cudaMemcpyHostToDevice, st));
    CKD(cudaMemcpyAsync(d_id_chars, id_chars.data(), id_chars.size(), cudaMemcpyHostToDevice, st));
    CKD(cudaMemcpyAsync(d_id_start, id_start.data(), (n_seq + 1) * 4, cudaMemcpyHostToDevice, st));
    sq.chars = dev_seq_chars; sq.start = d_start; sq.id_chars = d_id_chars; sq.id_start = d_id_start; sq.n = (uint32_t)n_seq;

    uint32_t* d_bad = nullptr;
    CKD(sc.alloc(&d_bad, 2));
    CKD(cudaMemsetAsync(d_bad, 0, 8, st));
    k_gfa_check<<<grid_for(m), 256, 0, st>>>(R.chr, R.pos, R.id, m, sq, k, d_bad);
    uint32_t bad = 0;
    CKD(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CKD(cudaStreamSynchronize(st));
    if (bad & 1u) return set_error("The input is corrupted");
    if (bad & 2u) return set_error("A vertex id is too large, cannot generate GFA");

    long long* sid = nullptr;
    uint32_t *uniq = nullptr, *uniq_before = nullptr;
    unsigned long long* key = nullptr;
    uint8_t* first = nullptr;
    CKD(sc.alloc(&sid, m));
    CKD(sc.alloc(&uniq, m));
    CKD(sc.alloc(&uniq_before, m));
    CKD(sc.alloc(&key, m));
    CKD(sc.alloc(&first, m));
    k_gfa_segment_ids<<<grid_for(m), 256, 0, st>>>(R.chr, R.pos, R.id, m, sq, k, sid, uniq);
    if (int rc = exclusive_sum(sc, uniq, uniq_before, m)) return rc;
    k_gfa_reserved_ids<<<grid_for(m), 256, 0, st>>>(uniq, uniq_before, m, sid);
    k_gfa_keys<<<grid_for(m), 256, 0, st>>>(R.chr, sid, m, key);
    Classes C;
    if (int rc = build_classes(sc, key, m, &C)) return rc;
    k_gfa_first<<<grid_for(m), 256, 0, st>>>(C.idx, C.head, C.key_sorted, m, first);

    uint32_t *chr_first = nullptr, *chr_last = nullptr;
    CKD(sc.alloc(&chr_first, n_seq));
    CKD(sc.alloc(&chr_last, n_seq));
    k_gfa_chr_bounds<<<grid_for(m), 256, 0, st>>>(R.chr, m, chr_first, chr_last);

    SegCtx ctx{R.chr, R.pos, sid, first, sq, k, format};
    unsigned long long *len_a = nullptr, *piece = nullptr, *piece_off = nullptr, *len = nullptr, *off = nullptr;
    CKD(sc.alloc(&len_a, m));
    CKD(sc.alloc(&piece, m + 1));
    CKD(sc.alloc(&piece_off, m + 1));
    CKD(sc.alloc(&len, m + 1));
    CKD(sc.alloc(&off, m + 1));
    CKD(cudaMemsetAsync(piece + m, 0, 8, st));
    CKD(cudaMemsetAsync(len + m, 0, 8, st));
    k_gfa_len<<<grid_for(m), 256, 0, st>>>(ctx, m, len_a, piece);
    if (int rc = exclusive_sum(sc, piece, piece_off, m + 1)) return rc;
    k_gfa_len_total<<<grid_for(m), 256, 0, st>>>(ctx, m, len_a, piece_off, chr_first, len);
    if (int rc = exclusive_sum(sc, len, off, m + 1)) return rc;
    unsigned long long total = 0;
    CKD(cudaMemcpyAsync(&total, off + m, 8, cudaMemcpyDeviceToHost, st));
    CKD(cudaStreamSynchronize(st));

    // long bodies: counted first (cap 0), then listed -- any number of them
    uint32_t *long_list = nullptr, *long_count = nullptr;
    CKD(sc.alloc(&long_count, 1));
    CKD(cudaMemsetAsync(long_count, 0, 4, st));
    k_gfa_long_list<<<grid_for(m), 256, 0, st>>>(ctx, m, nullptr, long_count, 0);
    uint32_t n_long = 0;
    CKD(cudaMemcpyAsync(&n_long, long_count, 4, cudaMemcpyDeviceToHost, st));
    CKD(cudaStreamSynchronize(st));
    CKD(sc.alloc(&long_list, n_long));
    char* text = nullptr;
    CKD(cudaMallocAsync((void**)&text, std::max<unsigned long long>(total, 16), st));
    k_gfa_write<<<grid_for(m), 256, 0, st>>>(ctx, m, off, len_a, piece_off, chr_first, chr_last, text);
    k_gfa_bodies_warp<<<grid_for(m), 256, 0, st>>>(ctx, m, off, text);
    cudaError_t e = cudaSuccess;
    if (n_long) {
        e = cudaMemsetAsync(long_count, 0, 4, st);
        k_gfa_long_list<<<grid_for(m), 256, 0, st>>>(ctx, m, long_list, long_count, n_long);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        k_gfa_bodies_cta<<<sms * 4, 256, 0, st>>>(ctx, long_list, n_long, off, text);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFreeAsync(text, st);
        return set_error("CUDA error %s in graphdump (%s)", cudaGetErrorName(e), cudaGetErrorString(e));
    }
    *dev_text = (uint8_t*)text;
    *text_bytes = total;
    return 0;
}

// what `graphdump -f gfa1|gfa2|fasta -k <k> [-s <fasta>]... [--prefix] <image>` prints (out_path NULL or "-" = stdout)
int tpc_graphdump_gfa_file(const char* image_path, const char* format, uint32_t k, const char* const* seq_paths, size_t n_seq_paths,
                           int prefix, const char* out_path) {
    if (!image_path || !format || (n_seq_paths && !seq_paths)) return set_error("null argument");
    const uint32_t fmt = !strcmp(format, "gfa1") ? kGfa1 : !strcmp(format, "gfa2") ? kGfa2 : !strcmp(format, "fasta") ? kFasta : 0u;
    if (!fmt) return set_error("format must be gfa1, gfa2 or fasta");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error("no CUDA device: twopaco_b200 has no CPU fallback");
    HostSequences hs;
    std::string current_header;
    for (size_t i = 0; i < n_seq_paths; ++i)
        if (int rc = read_sequences(seq_paths[i], &hs, &current_header)) return rc;
    const uint64_t n_seq = hs.header.size();
    // segment names (ReadInputSequences, graphdump.cpp:183-196; the "s<n>_" counter is never advanced there) and the head
    // of the output (Header + ListInputSequences: gfa1 lists every sequence with the LAST file that holds its name)
    std::vector<std::string> name(n_seq);
    std::map<std::string, std::string> file_of;
    for (uint64_t c = 0; c < n_seq; ++c) {
        name[c] = (prefix && fmt != kFasta) ? "s0_" + hs.header[c] : hs.header[c];
        file_of[name[c]] = hs.file[c];
    }
    std::string head;
    if (fmt == kGfa1) {
        head = "H\tVN:Z:1.0\n";
        for (uint64_t c = 0; c < n_seq; ++c) head += "S\t" + name[c] + "\t*\tUR:Z:" + file_of[name[c]] + "\n";
    } else if (fmt == kGfa2) head = "H\tVN:Z:2.0\n";
    std::vector<const char*> name_ptr(n_seq);
    for (uint64_t c = 0; c < n_seq; ++c) name_ptr[c] = name[c].c_str();

    FILE* f = fopen(image_path, "rb");
    if (!f) return set_error("Can't open file %s", image_path);
    fseek(f, 0, SEEK_END);
    const uint64_t bytes = (uint64_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> image(bytes);
    const bool read_ok = bytes == 0 || fread(image.data(), 1, bytes, f) == bytes;
    fclose(f);
    if (!read_ok) return set_error("Can't read file %s", image_path);

    uint8_t *d_img = nullptr, *d_seq = nullptr, *d_text = nullptr;
    uint64_t tbytes = 0;
    int rc = 0;
    if (cudaMalloc(&d_img, std::max<uint64_t>(bytes, 16)) != cudaSuccess || cudaMalloc(&d_seq, std::max<uint64_t>(hs.chars.size(), 16)) != cudaSuccess)
        rc = set_error("out of device memory for the image and the sequences");
    if (rc == 0 && bytes && cudaMemcpy(d_img, image.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = set_error("host to device copy failed");
    if (rc == 0 && !hs.chars.empty() && cudaMemcpy(d_seq, hs.chars.data(), hs.chars.size(), cudaMemcpyHostToDevice) != cudaSuccess)
        rc = set_error("host to device copy failed");
    if (rc == 0) {
        std::vector<uint64_t> start(hs.start.begin(), hs.start.end());
        rc = tpc_graphdump_gfa_device(d_img, bytes, fmt, k, d_seq, start.data(), name_ptr.data(), n_seq, nullptr, &d_text, &tbytes);
    }
    if (rc == 0) {
        FILE* o = (!out_path || !strcmp(out_path, "-")) ? stdout : fopen(out_path, "wb");
        if (!o) rc = set_error("Can't create the output file");
        if (rc == 0 && !head.empty() && fwrite(head.data(), 1, head.size(), o) != head.size()) rc = set_error("Can't write to the output file");
        const uint64_t kPiece = 64ull << 20;
        std::vector<uint8_t> piece((size_t)std::min<uint64_t>(std::max<uint64_t>(tbytes, 1), kPiece));
        for (uint64_t lo = 0; rc == 0 && lo < tbytes; lo += kPiece) {
            const uint64_t n = std::min(kPiece, tbytes - lo);
            if (cudaMemcpy(piece.data(), d_text + lo, n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_error("device to host copy failed");
            else if (fwrite(piece.data(), 1, n, o) != n) rc = set_error("Can't write to the output file");
        }
        if (o && o != stdout && fclose(o) != 0 && rc == 0) rc = set_error("Can't write to the output file");
        if (o == stdout) fflush(stdout);
    }
    if (d_text) cudaFree(d_text);
    if (d_img) cudaFree(d_img);
    if (d_seq) cudaFree(d_seq);
    return rc;
}

}  // extern "C"
