// tpc_host.cpp -- host side of libtwopaco_b200.so: FASTA framing, 2-bit packing and the level-1
// drop-in entry point tpc_build() (== TwoPaCo::CreateEnumerator, vertexenumerator.h:37-46).
// No CUDA kernels here; all compute goes through the session API (tpc_session.cu).
#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include <fcntl.h>
#include <unistd.h>

#include "tpc_ingest.h"
#include "tpc_internal.h"
#include "tpc_multi.h"

using tpc::set_error;

namespace {

// dnachar.cpp:9-11: VALID_CHARS; :18-33 MakeUpChar.  Table value: 0..3 = ACGT, 4 = valid
// non-definite (-> 'N', vertexenumerator.h:1174), 5 = whitespace, 6 = invalid.
struct CharTable {
    uint8_t t[256];
    CharTable() {
        for (int i = 0; i < 256; ++i) t[i] = isspace(i) ? 5 : 6;
        for (const char* p = "ACGTURYKMSWBDHWNXV"; *p; ++p) {
            t[(unsigned char)*p] = 4;
            t[(unsigned char)tolower(*p)] = 4;  // GetChar upper-cases first (streamfastaparser.cpp:79-88)
        }
        const char* acgt = "ACGT";
        for (int i = 0; i < 4; ++i) {
            t[(unsigned char)acgt[i]] = (uint8_t)i;
            t[(unsigned char)tolower(acgt[i])] = (uint8_t)i;
        }
    }
};
const CharTable kChars;

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// FASTA framing (streamfastaparser.cpp:29-133)
// ---------------------------------------------------------------------------------------------
int tpc_read_fasta(const char* path, char*** records, uint64_t** rec_len, uint64_t* n_records) {
    if (!path || !records || !rec_len || !n_records) return set_error("null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return set_error("Can't open file %s", path);
    std::vector<char> buf(1 << 20);  // the reference also reads 1 MiB blocks (streamfastaparser.h BUF_SIZE)
    uint64_t n = *n_records, cap = n;
    char** recs = *records;
    uint64_t* lens = *rec_len;
    std::string cur;
    std::string header;
    enum { kStart, kHeader, kSeq } state = kStart;
    bool have_record = false;
    int rc = 0;
    auto flush = [&]() -> int {
        if (!have_record) return 0;
        if (n == cap) {
            cap = cap ? cap * 2 : 16;
            char** nr = (char**)realloc(recs, cap * sizeof(char*));
            uint64_t* nl = (uint64_t*)realloc(lens, cap * sizeof(uint64_t));
            if (!nr || !nl) return set_error("out of memory");
            recs = nr; lens = nl;
        }
        char* s = (char*)malloc(cur.size() + 1);
        if (!s) return set_error("out of memory");
        memcpy(s, cur.data(), cur.size());
        s[cur.size()] = 0;
        recs[n] = s; lens[n] = cur.size(); ++n;
        cur.clear();
        have_record = false;
        return 0;
    };
    size_t got;
    while (rc == 0 && (got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < got && rc == 0; ++i) {
            unsigned char ch = (unsigned char)buf[i];
            switch (state) {
                case kStart:
                    if (ch != '>') { rc = set_error("The FASTA header should start with a '>', started with '%c'", ch); break; }
                    state = kHeader; header.clear(); have_record = true;
                    break;
                case kHeader:
                    if (ch == '\n') state = kSeq; else header.push_back((char)ch);
                    break;
                case kSeq: {
                    uint8_t c = kChars.t[ch];
                    if (c < 4) cur.push_back("ACGT"[c]);
                    else if (c == 4) cur.push_back('N');
                    else if (c == 5) {}
                    else if (ch == '>') { rc = flush(); state = kHeader; header.clear(); have_record = true; }
                    else {
                        std::istringstream hs(header); std::string first; hs >> first;
                        rc = set_error("Found an invalid character '%c' in sequence %s", ch, first.c_str());
                    }
                    break;
                }
            }
        }
    }
    fclose(f);
    if (rc == 0) rc = flush();
    *records = recs; *rec_len = lens; *n_records = n;
    return rc;
}

void tpc_free_records(char** records, uint64_t* rec_len, uint64_t n_records) {
    if (records) for (uint64_t i = 0; i < n_records; ++i) free(records[i]);
    free(records);
    free(rec_len);
}

// ---------------------------------------------------------------------------------------------
// 2-bit packing of the whole input (layout: include/twopaco_b200.h, tpc_genome)
// ---------------------------------------------------------------------------------------------
int tpc_pack_records(const char* const* records, const uint64_t* rec_len, uint64_t n_records, uint32_t threads,
                     uint64_t* codes, uint64_t* n_mask, uint64_t* rec_start) {
    if ((n_records && (!records || !rec_len)) || !codes || !n_mask) return set_error("null argument");
    uint64_t npos = tpc_positions_for(rec_len, n_records);
    uint64_t cw = tpc_code_words(npos), mw = tpc_mask_words(npos);
    std::vector<uint64_t> start(n_records);
    uint64_t p = 1;
    for (uint64_t i = 0; i < n_records; ++i) { start[i] = p; p += rec_len[i] + 1; }
    if (rec_start) std::copy(start.begin(), start.end(), rec_start);

    // Work unit = 64 positions (one n_mask word, two code words), so threads never share words.
    const uint64_t blocks = mw;  // covers the padding too
    threads = std::max<uint32_t>(1, std::min<uint32_t>(threads, 256));
    std::atomic<uint64_t> next{0};
    const uint64_t chunk = 1 << 12;  // 4096 blocks = 256 Ki positions per grab
    auto worker = [&]() {
        for (;;) {
            uint64_t b0 = next.fetch_add(chunk);
            if (b0 >= blocks) break;
            uint64_t b1 = std::min(blocks, b0 + chunk);
            // record containing (or following) the first position of the chunk
            uint64_t pos = b0 * 64;
            uint64_t r = std::upper_bound(start.begin(), start.end(), pos) - start.begin();
            r = r ? r - 1 : 0;
            for (uint64_t b = b0; b < b1; ++b) {
                uint64_t lo = 0, hi = 0, nm = 0;
                for (uint32_t j = 0; j < 64; ++j) {
                    uint64_t q = b * 64 + j;
                    uint64_t code = 0, isn = 1;
                    if (q < npos && n_records) {
                        while (r + 1 < n_records && q >= start[r + 1]) ++r;
                        if (q >= start[r] && q < start[r] + rec_len[r]) {
                            uint8_t c = kChars.t[(unsigned char)records[r][q - start[r]]];
                            if (c < 4) { code = c; isn = 0; }
                        }
                    }
                    if (j < 32) lo |= code << (2 * j); else hi |= code << (2 * (j - 32));
                    nm |= isn << j;
                }
                if (2 * b < cw) codes[2 * b] = lo;
                if (2 * b + 1 < cw) codes[2 * b + 1] = hi;
                n_mask[b] = nm;
            }
        }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// level 1: tpc_build == CreateEnumerator
// ---------------------------------------------------------------------------------------------
namespace {

// IngestSink that ships every span to the device: two pinned staging buffers, the copy of span i
// overlaps the parsing of span i+1.
class StagedUpload : public tpc::IngestSink {
public:
    static constexpr uint64_t kSpan = 64ull << 20;
    explicit StagedUpload(uint8_t* device_base) : dev_(device_base) {}
    ~StagedUpload() override {
        for (int i = 0; i < 2; ++i) {
            if (buf_[i]) cudaFreeHost(buf_[i]);
            if (ev_[i]) cudaEventDestroy(ev_[i]);
        }
    }
    uint8_t* acquire(uint64_t max_bytes) override {
        int i = next_;
        if (ev_[i] && cudaEventSynchronize(ev_[i]) != cudaSuccess) return nullptr;   // previous copy out of this buffer done
        if (cap_[i] < max_bytes) {
            if (buf_[i]) cudaFreeHost(buf_[i]);
            buf_[i] = nullptr;
            cap_[i] = (max_bytes + (1u << 20) - 1) >> 20 << 20;   // small inputs pin little
            if (cudaMallocHost((void**)&buf_[i], cap_[i]) != cudaSuccess) { cap_[i] = 0; return nullptr; }
        }
        return buf_[i];
    }
    int commit(uint64_t off, uint8_t* buf, uint64_t nbytes) override {
        int i = next_;
        if (cudaMemcpyAsync(dev_ + off, buf, nbytes, cudaMemcpyHostToDevice, nullptr) != cudaSuccess)
            return set_error("host to device copy failed");
        if (!ev_[i]) cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming);
        cudaEventRecord(ev_[i], nullptr);
        next_ ^= 1;
        return 0;
    }
    int finish() { return cudaDeviceSynchronize() == cudaSuccess ? 0 : set_error("host to device copy failed"); }

private:
    uint8_t* dev_;
    uint8_t* buf_[2] = {nullptr, nullptr};
    uint64_t cap_[2] = {0, 0};
    cudaEvent_t ev_[2] = {nullptr, nullptr};
    int next_ = 0;
};

int write_chunk_to_file(void* ctx, const uint8_t* data, uint64_t nbytes) {
    return fwrite(data, 1, nbytes, (FILE*)ctx) == nbytes ? 0 : set_error("Can't write to the output file");
}

}  // namespace

struct tpc_handle {
    tpc_session* session = nullptr;
    void *d_codes = nullptr, *d_nmask = nullptr;
    tpc_stats stats{};
    uint64_t junctions = 0;
    int device = 0;          // the GPU that holds the session and the genome (GetId)
    uint32_t gpus = 1;
};

namespace {
struct Logger {
    tpc_log_fn fn; void* ctx;
    void operator()(const std::string& s) const { if (fn) fn(ctx, s.c_str()); }
};
}  // namespace

extern "C" {

int tpc_build(const char* const* fasta_paths, size_t n_files, uint32_t k, uint32_t filter_bits, uint32_t q,
              uint32_t rounds, uint32_t threads, uint64_t abundance, const char* tmpdir, const char* outfile,
              tpc_log_fn log, void* log_ctx, tpc_handle** out) {
    (void)tmpdir;  // no temp files: candidate masks and junction keys stay in HBM (reference: h:219, 292)
    if (!fasta_paths || !outfile || !out) return set_error("null argument");
    *out = nullptr;
    Logger L{log, log_ctx};
    tpc_params prm{};
    prm.k = k; prm.filter_bits = filter_bits; prm.q = q; prm.rounds = rounds ? rounds : 1;
    prm.abundance = abundance; prm.shard_index = 0; prm.shard_count = 1; prm.seed = 0;
    tpc_session* s = nullptr;
    if (int rc = tpc_session_create(&prm, nullptr, &s)) return rc;

    {   // log header, same lines as vertexenumerator.h:137-147
        std::ostringstream ss;
        ss << "Threads = " << threads << "\nVertex length = " << k << "\nHash functions = " << q
           << "\nFilter size = " << (1ull << filter_bits) << "\nCapacity = " << (k + 4 + 31) / 32 << "\nFiles: \n";
        for (size_t i = 0; i < n_files; ++i) ss << fasta_paths[i] << "\n";
        L(ss.str());
    }

    // FASTA -> position layout (all host threads, parsed once), streamed to the device through two
    // pinned staging buffers, packed to 2 bits + N mask by K0 on the GPU
    const bool verbose = getenv("TPC_VERBOSE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    tpc::IngestPlan plan;
    int rc = tpc::ingest_plan(fasta_paths, n_files, threads ? threads : 1, &plan);
    double t_plan = now();
    void *d_ascii = nullptr, *d_codes = nullptr, *d_nmask = nullptr;
    tpc_genome g{};
    if (rc == 0) rc = tpc_device_alloc(plan.layout_bytes, &d_ascii);
    if (rc == 0) rc = tpc_device_alloc(tpc_code_words(plan.n_positions) * 8, &d_codes);
    if (rc == 0) rc = tpc_device_alloc(tpc_mask_words(plan.n_positions) * 8, &d_nmask);
    if (rc == 0) {
        StagedUpload up((uint8_t*)d_ascii);
        rc = tpc::ingest_emit(plan, threads ? threads : 1, StagedUpload::kSpan, up);
        if (rc == 0) rc = up.finish();
    }
    if (rc == 0) rc = tpc_pack_ascii_device((const uint8_t*)d_ascii, plan.n_positions, (uint64_t*)d_codes, (uint64_t*)d_nmask, nullptr);
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) rc = set_error("K0 pack failed");
    tpc_device_free(d_ascii);
    if (rc == 0) {
        g.codes = (const uint64_t*)d_codes; g.n_mask = (const uint64_t*)d_nmask; g.n_positions = plan.n_positions;
        g.rec_start = plan.rec_start.data(); g.rec_len = plan.rec_len.data(); g.n_records = plan.rec_start.size();
        rc = tpc_session_set_genome_device(s, &g);
    }
    double t_upload = now();
    uint64_t bytes = 0;
    // How many GPUs: every visible device when the input is worth sharding, TPC_GPUS=<n> overrides (SURVEY 8(b):
    // "-t keeps meaning host worker threads, GPU count from a new env/flag or CUDA_VISIBLE_DEVICES")
    uint32_t gpus = 1;
    if (rc == 0) {
        const uint32_t visible = tpc_visible_gpus();
        // creating the NCCL communicators takes seconds (measured: 5.6 s for 2 GPUs), more than one GPU needs for 20 Gbp, so
        // several GPUs are opt-in (TPC_GPUS=<n>, 0 = all visible) unless one GPU cannot hold the input at all
        if (const char* e = getenv("TPC_GPUS")) gpus = atoi(e) > 0 ? (uint32_t)atoi(e) : visible;
        else if (plan.n_positions >= (1ull << 37)) gpus = visible;
        gpus = std::max<uint32_t>(1, std::min(gpus, visible));
        int cur = 0;
        cudaGetDevice(&cur);
        if (gpus > 1 && cur != 0) gpus = 1;   // the genome was packed on the current device; shards start at device 0
    }
    tpc_stats st{};
    tpc_multi* multi = nullptr;
    if (rc == 0 && gpus > 1) {
        // N GPUs of this process: chunked NCCL broadcast of the packed genome from GPU 0, hash-range shards, NCCL
        // exchanges between the stages, every GPU pwrite()s its slice of the image (tpc_multi.cpp)
        tpc_session_destroy(s);
        s = nullptr;
        rc = tpc_multi_create(gpus, nullptr, &multi);
        if (rc == 0) {
            int fd = open(outfile, O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (fd < 0) rc = set_error("Can't create the output file");
            else {
                rc = tpc::multi_run_from_device0(multi, &prm, (const uint64_t*)d_codes, (const uint64_t*)d_nmask, plan.n_positions,
                                                 plan.rec_start.data(), plan.rec_len.data(), plan.rec_start.size(), fd, &bytes, &st, &s);
                if (close(fd) != 0 && rc == 0) rc = set_error("Can't write to the output file");
            }
        }
        if (multi) tpc_multi_destroy(multi);
    } else {
        if (rc == 0) rc = tpc_session_run_to_count(s, &bytes);
    }
    double t_gpu = now();
    if (rc == 0 && gpus == 1) {
        // JunctionPositionWriter (junctionapi.h:110-116): creates / truncates the output file; the image
        // is produced on the device and streamed to the file through two pinned staging buffers
        FILE* f = fopen(outfile, "wb");
        if (!f) rc = set_error("Can't create the output file");
        else {
            rc = tpc_session_write_stream(s, bytes, write_chunk_to_file, f);
            if (fclose(f) != 0 && rc == 0) rc = set_error("Can't write to the output file");
        }
    }
    if (rc == 0 && gpus == 1) rc = tpc_session_stats(s, &st);
    if (verbose)
        fprintf(stderr, "[tpc_build] %u GPU(s): frame+count %.3f s, normalise+upload+pack %.3f s, gpu passes %.3f s, emit+write %.3f s\n",
                gpus, t_plan - t0, t_upload - t_plan, t_gpu - t_upload, now() - t_gpu);
    if (rc == 0) {
        std::ostringstream ss;
        ss << std::string(80, '-') << "\n"
           << "Round 0, 0:" << (1ull << filter_bits) << "\nPass\tFilling\tFiltering\n"
           << "1\t" << (int)(st.ms_fill / 1000) << "\t" << (int)(st.ms_query / 1000) << "\t\n"
           << "2\t" << (int)(st.ms_insert / 1000) << "\t" << (int)(st.ms_classify / 1000) << "\n"
           << "True junctions count = " << st.junctions << "\n"
           << "False junctions count = " << (st.candidate_kmers - st.junctions) << "\n"
           << "Hash table size = " << st.candidate_kmers << "\n"
           << "Candidate marks count = " << st.candidate_marks << "\n"
           << std::string(80, '-') << "\n"
           << "Reallocating bifurcations time: " << (int)(st.ms_index / 1000) << "\n"
           << "True marks count: " << st.occurrences << "\n"
           << "Edges construction time: " << (int)(st.ms_emit / 1000) << "\n"
           << std::string(80, '-') << "\n";
        L(ss.str());
    }
    tpc_handle* h = rc == 0 ? new (std::nothrow) tpc_handle() : nullptr;
    if (!h) {
        if (s) tpc_session_destroy(s);
        tpc_device_free(d_codes);
        tpc_device_free(d_nmask);
        return rc ? rc : set_error("out of memory");
    }
    h->session = s; h->stats = st; h->junctions = st.junctions; h->gpus = gpus;
    cudaGetDevice(&h->device);
    h->d_codes = d_codes; h->d_nmask = d_nmask;  // the session reads k-mers back from the genome (GetId)
    *out = h;
    return 0;
}

uint64_t tpc_vertices(const tpc_handle* h) { return h ? h->junctions : 0; }

int64_t tpc_get_id(const tpc_handle* h, const char* kmer) {
    int64_t id = TPC_INVALID_VERTEX;
    if (!h || !h->session) return id;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != h->device) cudaSetDevice(h->device);
    const int rc = tpc_session_get_id(h->session, kmer, &id);
    if (cur != h->device) cudaSetDevice(cur);
    return rc != 0 ? TPC_INVALID_VERTEX : id;
}

int tpc_handle_stats(const tpc_handle* h, tpc_stats* out) {
    if (!h || !out) return set_error("null argument");
    *out = h->stats;
    return 0;
}

void tpc_free(tpc_handle* h) {
    if (!h) return;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != h->device) cudaSetDevice(h->device);
    tpc_session_destroy(h->session);
    tpc_device_free(h->d_codes);
    tpc_device_free(h->d_nmask);
    if (cur != h->device) cudaSetDevice(cur);
    delete h;
}

}  // extern "C"
