// tpc_session.cu -- sessions (C ABI level 3), the packed-genome entry point (level 2) and the
// W-independent kernels.  One session = one GPU = one hash-range shard.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/twopaco_b200.h"
#include "tpc_internal.h"
#include "tpc_kernels_common.cuh"
#include "tpc_launch.cuh"
#include "tpc_window_provider.h"

namespace tpc {

// ---------------------------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_error;

int set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}
const char* last_error() { return g_error.c_str(); }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return tpc::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                                  cudaGetErrorString(e_));                                         \
    } while (0)

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// TPC_VERBOSE: host-side trace points with one clock for the whole process (where does the time outside the kernels go?)
static void trace(const char* what) {
    static const bool on = getenv("TPC_VERBOSE") != nullptr;
    static const double t0 = now_ms();
    if (on) fprintf(stderr, "[tpc trace %d] %12.3f ms  %s\n", (int)getpid(), now_ms() - t0, what);
}
// host wall clock of a scope, added to *acc on exit (also on the early returns of CK())
struct WallTimer {
    float* acc;
    double t0;
    explicit WallTimer(float* a) : acc(a), t0(now_ms()) {}
    ~WallTimer() { *acc += (float)(now_ms() - t0); }
};

// RAII holders so that the early returns of CK() do not leak events / buffers
struct Events {
    std::vector<cudaEvent_t> e;
    explicit Events(int n, unsigned flags = cudaEventDefault) : e(n, nullptr) {
        for (auto& x : e) cudaEventCreateWithFlags(&x, flags);
    }
    ~Events() {
        for (auto x : e)
            if (x) cudaEventDestroy(x);
    }
    cudaEvent_t operator[](int i) const { return e[i]; }
    Events(const Events&) = delete;
    Events& operator=(const Events&) = delete;
};
struct DevBuf {   // stream-ordered device allocation released on scope exit
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ~DevBuf() {
        if (p) cudaFreeAsync(p, st);
    }
};
struct PinnedBuf {
    void* p = nullptr;
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
};

// ---------------------------------------------------------------------------------------------
// device memory: stream-ordered allocations from the device's default pool, which is told to keep
// freed memory (release threshold = max) so that repeated runs do not pay cudaMalloc/cudaFree.
// ---------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t dev_alloc(T** p, size_t bytes, cudaStream_t st) {
    return cudaMallocAsync(reinterpret_cast<void**>(p), bytes ? bytes : 16, st);
}
static cudaError_t dev_free(void* p, cudaStream_t st) { return p ? cudaFreeAsync(p, st) : cudaSuccess; }

static void configure_pool(int device) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
}

// memory a new allocation can use: free device memory + what the pool holds but does not use.
// cudaMemGetInfo goes through the driver and was measured (B200 box shared with other tenants, profiles/r2_host_trace.md) to
// block for 30-100 ms every few calls; it sat three times on the critical path of every find_candidates call.  What this
// process can use in total -- free memory + what its pool has reserved -- only changes when somebody else (another library
// of the process, another process) allocates or frees, so that sum is cached per device: refreshed when it is older than
// a minute or after an allocation failed (invalidate_memory_budget), and what the pool currently uses (a counter of the
// pool, no driver call) is subtracted.
namespace {
struct MemoryBudget { double measured_ms = -1e30; uint64_t usable = 0, total = 0; };
MemoryBudget g_budget[64];
std::mutex g_budget_mu;
}  // namespace
static void invalidate_memory_budget() {
    std::lock_guard<std::mutex> lock(g_budget_mu);
    for (MemoryBudget& b : g_budget) b.measured_ms = -1e30;
}
static uint64_t available_bytes(int device, uint64_t* total_out = nullptr) {
    cudaMemPool_t pool;
    uint64_t reserved = 0, used = 0;
    const bool have_pool = cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess;
    if (have_pool) cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
    std::lock_guard<std::mutex> lock(g_budget_mu);
    MemoryBudget& b = g_budget[(unsigned)device % 64];
    if (now_ms() - b.measured_ms > 60000.0) {   // (a stale value costs at most a retry: a failed allocation invalidates it)
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (have_pool) cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        b.usable = (uint64_t)free_b + reserved;
        b.total = total_b;
        b.measured_ms = now_ms();
    }
    if (total_out) *total_out = b.total;
    return b.usable > used ? b.usable - used : 0;
}

// ---------------------------------------------------------------------------------------------
// W-independent launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_classify(const LaunchCtx& c, TableView T, uint64_t abundance, uint32_t use_abundance,
                            unsigned long long* out, unsigned long long* out_keys, uint64_t out_cap, Counters* ctr) {
    uint64_t cap = 1ull << T.log2cap;
    uint64_t blocks = std::min<uint64_t>((cap + 255) / 256, (uint64_t)c.sm_count * 8);
    k_classify<<<(int)blocks, 256, 0, c.stream>>>(T, abundance, use_abundance, out, out_keys, out_cap, ctr);
    ++*c.launches;
    return cudaGetLastError();
}

uint64_t scan_scratch_items(uint64_t n) {
    uint64_t total = 0;
    while (n > (uint64_t)kScanBlock) {
        n = (n + kScanBlock - 1) / kScanBlock;
        total += n;
    }
    return total + 1;
}

// in-place exclusive scan; scratch holds scan_scratch_items(n) words
cudaError_t launch_scan_exclusive(const LaunchCtx& c, unsigned long long* data, uint64_t n, unsigned long long* scratch) {
    if (n == 0) return cudaSuccess;
    uint64_t nb = (n + kScanBlock - 1) / kScanBlock;
    if (nb == 1) {
        k_scan_apply<<<1, 256, 0, c.stream>>>(data, n, nullptr);
        ++*c.launches;
        return cudaGetLastError();
    }
    k_scan_reduce<<<(unsigned)nb, 256, 0, c.stream>>>(data, n, scratch);
    ++*c.launches;
    cudaError_t e = launch_scan_exclusive(c, scratch, nb, scratch + nb);
    if (e != cudaSuccess) return e;
    k_scan_apply<<<(unsigned)nb, 256, 0, c.stream>>>(data, n, scratch);
    ++*c.launches;
    return cudaGetLastError();
}

static int apply_grid(const LaunchCtx& c) { return c.sm_count * (c.apply_ctas > 0 && c.apply_ctas < 4 ? c.apply_ctas : 4); }

#define TPC_APPLY_Q_SWITCH(q, ...)                             \
    switch (q) {                                               \
        case 1: { constexpr int Q = 1; __VA_ARGS__; } break;   \
        case 2: { constexpr int Q = 2; __VA_ARGS__; } break;   \
        case 3: { constexpr int Q = 3; __VA_ARGS__; } break;   \
        case 4: { constexpr int Q = 4; __VA_ARGS__; } break;   \
        case 5: { constexpr int Q = 5; __VA_ARGS__; } break;   \
        case 6: { constexpr int Q = 6; __VA_ARGS__; } break;   \
        case 7: { constexpr int Q = 7; __VA_ARGS__; } break;   \
        default: { constexpr int Q = 8; __VA_ARGS__; } break;  \
    }

// where slice `bucket`'s records live: uniform layout, or the per-slice one of a re-binned skewed round (host copies)
struct SliceLayout {
    const std::vector<unsigned long long>* off = nullptr;
    const std::vector<unsigned long long>* capv = nullptr;
    const uint32_t* base(const BinView& bv, uint32_t b) const { return bv.rec + 3 * (off ? (*off)[b] : (uint64_t)b * bv.cap); }
    uint64_t cap(const BinView& bv, uint32_t b) const { return capv ? (*capv)[b] : bv.cap; }
};

cudaError_t launch_apply_fill(const LaunchCtx& c, uint32_t* filter, const BinView& bv, const SliceLayout& sl, uint32_t bucket, Counters* ctr) {
    uint32_t* slice = filter + (((uint64_t)bucket << bv.sib_bits) << 3);
    TPC_APPLY_Q_SWITCH(bv.q, (k_apply_fill<Q><<<apply_grid(c), 256, 0, c.stream>>>(
        slice, sl.base(bv, bucket), bv.count + bucket, sl.cap(bv, bucket), (1u << bv.sib_bits) - 1u, ctr)));
    ++*c.launches;
    return cudaGetLastError();
}

constexpr bool kQueryAggDefault = true;   // measured at C3, one GPU, same box: query 117.4 -> 109.5 ms (profiles/r2_query_agg.md)

cudaError_t launch_apply_query(const LaunchCtx& c, const uint32_t* filter, const BinView& bv, const SliceLayout& sl, uint32_t bucket,
                               uint32_t* mask, uint64_t wave_base, Counters* ctr, uint32_t* hll, const MarkList& ml) {
    const uint32_t* slice = filter + (((uint64_t)bucket << bv.sib_bits) << 3);
    // (TPC_QUERY_AGG=0 / 1: marks handled inline by the lane that finds them / queued per warp and handled 32 at a time)
    const char* agg_env = getenv("TPC_QUERY_AGG");
    const bool agg = agg_env ? atoi(agg_env) != 0 : kQueryAggDefault;
    if (agg) {
        TPC_APPLY_Q_SWITCH(bv.q, (k_apply_query<Q, true><<<apply_grid(c), 256, 0, c.stream>>>(
            slice, sl.base(bv, bucket), bv.count + bucket, sl.cap(bv, bucket), bv.sib_bits, mask, wave_base, ctr, hll,
            (uint64_t)bucket << bv.sib_bits, ml)));
    } else {
        TPC_APPLY_Q_SWITCH(bv.q, (k_apply_query<Q, false><<<apply_grid(c), 256, 0, c.stream>>>(
            slice, sl.base(bv, bucket), bv.count + bucket, sl.cap(bv, bucket), bv.sib_bits, mask, wave_base, ctr, hll,
            (uint64_t)bucket << bv.sib_bits, ml)));
    }
    ++*c.launches;
    return cudaGetLastError();
}

cudaError_t launch_apply_overflow(const LaunchCtx& c, uint32_t* filter, const BinView& bv, int do_query, uint32_t* mask,
                                  uint64_t wave_base, Counters* ctr, uint32_t* hll, const MarkList& ml) {
    k_apply_overflow<<<c.sm_count, 256, 0, c.stream>>>(filter, bv.ov, bv.ov_count, bv.ov_cap, bv.sib_bits, bv.q, do_query, mask,
                                                       wave_base, ctr, hll, ml);
    ++*c.launches;
    return cudaGetLastError();
}

}  // namespace tpc

using namespace tpc;

// ---------------------------------------------------------------------------------------------
// session
// ---------------------------------------------------------------------------------------------
struct tpc_session {
    tpc_params prm{};
    cudaStream_t stream = nullptr;
    int device = 0;
    int sm_count = 148;
    int W = 1;
    uint32_t filter_bits_eff = 0;

    // genome
    GenomeView g{};
    uint64_t* d_codes = nullptr;   // owned copies (set_genome_host) or null
    uint64_t* d_nmask = nullptr;
    std::vector<uint64_t> rec_start, rec_len;
    std::vector<uint32_t> sep_before;
    std::vector<uint64_t> emit_prev;  // index of the last emitting record before record i (or 0)
    uint64_t* d_rec_start = nullptr;
    uint64_t* d_rec_len = nullptr;
    uint32_t* d_sep_before = nullptr;
    uint64_t ntiles = 0;

    // working set
    uint32_t* d_filter = nullptr;
    uint32_t* d_mask = nullptr;
    uint32_t* d_stubmask = nullptr;
    Slot* d_T = nullptr;
    uint64_t T_bytes = 0;
    uint32_t T_log2 = 0;
    Slot* d_J = nullptr;
    uint32_t J_log2 = 0;
    unsigned long long* d_local = nullptr;  // this shard's junction words
    uint64_t local_cap = 0, local_count = 0;
    unsigned long long* d_sorted = nullptr;  // all junction words, sorted by position
    uint64_t J_count = 0;
    void* d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    Counters* d_ctr = nullptr;
    uint32_t* d_hll = nullptr;     // HyperLogLog registers of the round's candidates
    long long* d_id = nullptr;

    // emit slice state
    unsigned long long* d_tile_rec = nullptr;
    unsigned long long* d_tile_stub = nullptr;
    unsigned long long* d_scan_scratch = nullptr;
    uint64_t tile_cap = 0;
    uint64_t slice_tile_begin = 0, slice_tile_end = 0, slice_pos_begin = 0, slice_pos_end = 0;
    uint64_t slice_records = 0, slice_stubs = 0;
    bool have_candidates = false, have_index = false, have_count = false;

    // binned filter passes (tpc_bin.cuh)
    int filter_mode = 0;           // 0 auto, 1 direct, 2 binned (env TPC_FILTER_MODE)
    uint32_t slice_log2 = 26;      // filter slice kept L2-resident by the apply kernels (env TPC_SLICE_LOG2)
    uint64_t bin_budget_bytes = 0; // 0 = 70 % of free HBM (env TPC_BIN_BUFFER_MB)
    uint32_t* d_bin_rec = nullptr;  // record scratch of the call: the rounds' record waves, then (between the
    uint64_t bin_rec_bytes = 0;     //   filter passes and the next round) the round's candidate table
    // The scratch is kept between find_candidates calls of a session when plenty of memory stays free beside it (the next call
    // then reuses it instead of asking the pool for tens of GB again, which can cost tens of ms when the pool has to remap)
    bool bin_keep = false;
    uint64_t bin_count_bytes = 0, bin_ov_bytes = 0, own_extra_bytes = 0;
    uint64_t kept_scratch_bytes() const { return d_bin_rec && !bin_ready ? bin_rec_bytes + bin_count_bytes + bin_ov_bytes : 0; }
    unsigned long long* d_bin_count = nullptr;
    uint32_t* d_bin_ov = nullptr;
    BinView bin_view{};             // layout of the scratch, fixed for all rounds of a call
    // per-slice layout of a re-binned skewed round (BinView::off / capv): host copies + device arrays
    std::vector<unsigned long long> skew_off, skew_cap;
    unsigned long long* d_skew = nullptr;   // [2][buckets]
    uint32_t skew_rebins = 0;
    uint64_t bin_wave_tiles = 0, bin_nwaves = 0;
    bool bin_ready = false;
    bool used_binned = false;
    bool T_in_scratch = false;
    // Pipelined rounds: the record scratch is split in two halves; while round r is FILLED from one half
    // (apply kernels: bound by L2 / L1-tag traffic, few instructions) round r+1 is BINNED into the other
    // on a second stream (k_bin_list: bound by the integer pipes, little memory traffic), each kernel
    // capped to a share of the SM so that both are resident.  The query of round r runs alone afterwards.
    int pipe_env = 0;               // env TPC_PIPELINE=1 enables (not the default: see choose_sub_rounds)
    // CTAs per SM while both run (the register file holds four 256-thread CTAs of these kernels).  Measured at C3
    // on one GPU: 2+2 gains nothing (both kernels also share the LSU / L1 data path: binning 355 ms instead of
    // 226 ms alone), 3+1 turns 640 ms per step into 593 ms.
    int pipe_bin_ctas = 3, pipe_fill_ctas = 1;   // env TPC_PIPE_BIN_CTAS / TPC_PIPE_FILL_CTAS
    bool pipe = false;
    cudaStream_t bin_stream = nullptr;
    BinView pipe_view[2]{};
    long long pipe_round[2] = {-1, -1};   // round whose records a half holds
    cudaEvent_t pipe_ev[6]{};       // [0,1] half binned  [2] main-stream marker  [3,4] bin start/stop (timing)

    // mark list (MarkList, tpc_kernels.cuh): positions of the candidates the binned query kernels found, appended per CTA;
    // the exact pass reads it instead of walking the whole mask when the marks are sparse
    unsigned long long* d_marklist = nullptr;
    uint32_t* d_marklist_counts = nullptr;
    uint32_t marklist_regions = 0, marklist_region_cap = 0;
    bool marklist_valid = false;    // every mark of the current insert group went through the binned query kernels
    MarkList mark_list(bool enabled) const {
        return MarkList{enabled ? d_marklist : nullptr, d_marklist_counts, marklist_regions, marklist_region_cap};
    }

    // position-windowed sessions (tpc_windowed.inl): the genome streams through HBM window by window from `wp`;
    // g / d_mask / d_stubmask then are VIRTUAL views of the current window's buffers
    uint64_t h2d_bytes = 0;   // bytes the session copied host -> device for the genome (set_genome_host)
    bool windowed = false;
    WindowProvider* wp = nullptr;
    uint64_t window_tiles = 0, w_npos = 0;
    uint32_t* d_wmask = nullptr;
    uint32_t* d_wstub = nullptr;
    unsigned long long* d_local_keys = nullptr;   // junction keys beside d_local (windowed runs)

    // hash sub-ranges processed in sequence by this GPU: the user's -r times the sub-rounds chosen so
    // that one round's records fit HBM in one wave (choose_sub_rounds); ownership planes of all of
    // them come from ONE scan (k_own), plane 0 lives in the stub mask until the emit stage
    uint32_t sub_rounds = 1, rounds_eff = 1;
    int sub_rounds_env = 0;        // env TPC_SUBROUNDS (0 = automatic)
    uint32_t* d_own_extra = nullptr;
    OwnPlanes own{};
    bool own_shared = false;
    uint64_t own_done_tiles = 0;

    // host->device upload of the genome overlapped with the first binning pass (set_genome_host)
    cudaStream_t copy_stream = nullptr;
    bool up_ev_owned = true;                 // false: events handed in by tpc_session_add_genome_event
    std::vector<cudaEvent_t> up_ev;          // upload chunk c is complete
    std::vector<uint64_t> up_tile_begin;     // first tile of chunk c
    size_t up_waited = 0;                    // chunks the compute stream already waits for

    tpc_stats st{};
    cudaEvent_t ev[12]{};
    uint32_t launches = 0;

    bool allow_inline = true;      // env TPC_INLINE_KEYS=0 forces position-identified slots for every k
    uint32_t inline_keys() const { return (allow_inline && W == 1 && prm.abundance == ~0ull) ? 1u : 0u; }
    int bin_ctas = 0, apply_ctas = 0;   // env TPC_BIN_CTAS / TPC_APPLY_CTAS (tuning aids)
    LaunchCtx lctx() { return LaunchCtx{stream, sm_count, &launches, bin_ctas, apply_ctas}; }
    KParams kparams(uint32_t part) const {
        KParams kp{};
        kp.k = prm.k;
        kp.q = std::min<uint32_t>(std::max<uint32_t>(prm.q, 1u), 8u);
        kp.sector_shift = 64 - (filter_bits_eff - 8);
        kp.nparts = rounds_eff * prm.shard_count;
        kp.part = part;
        kp.count_occurrences = prm.abundance != ~0ull;
        kp.seed = prm.seed;
        return kp;
    }
    RecordTable rtable() const { return RecordTable{d_rec_start, d_rec_len, d_sep_before, (uint64_t)rec_start.size()}; }
};

// one explicit instantiation of the kernels per k-mer word count (tpc_w1.cu .. tpc_w4.cu, tpc_wn.cu for 5..19 = k <= 603,
// the reference's MAX_CAPACITY, vertexenumerator.h:4 / vertexenumerator.cpp:17-54)
#define W_CASE(n, EXPR) (s_->W == n) ? Launch<n>::EXPR
#define W_DISPATCH(s, EXPR)                                                                                                   \
    ([&]() -> cudaError_t {                                                                                                   \
        const tpc_session* s_ = (s);                                                                                          \
        return W_CASE(1, EXPR) : W_CASE(2, EXPR) : W_CASE(3, EXPR) : W_CASE(4, EXPR) : W_CASE(5, EXPR) : W_CASE(6, EXPR)      \
             : W_CASE(7, EXPR) : W_CASE(8, EXPR) : W_CASE(9, EXPR) : W_CASE(10, EXPR) : W_CASE(11, EXPR) : W_CASE(12, EXPR)   \
             : W_CASE(13, EXPR) : W_CASE(14, EXPR) : W_CASE(15, EXPR) : W_CASE(16, EXPR) : W_CASE(17, EXPR)                   \
             : W_CASE(18, EXPR) : Launch<19>::EXPR;                                                                           \
    }())

static uint32_t ceil_log2(uint64_t x) {
    uint32_t l = 0;
    while ((1ull << l) < x) ++l;
    return l;
}

extern "C" {

const char* tpc_last_error(void) { return tpc::last_error(); }
uint32_t tpc_abi_version(void) { return TPC_ABI_VERSION; }

uint64_t tpc_code_words(uint64_t n_positions) {
    uint64_t tiles = (n_positions + kTilePos - 1) / kTilePos;
    return tiles * kTileThreads + kCodePadWords;
}
uint64_t tpc_mask_words(uint64_t n_positions) {
    uint64_t tiles = (n_positions + kTilePos - 1) / kTilePos;
    return tiles * (kTileThreads / 2) + kMaskPadWords;
}
uint64_t tpc_positions_for(const uint64_t* rec_len, uint64_t n_records) {
    uint64_t p = 1;
    for (uint64_t i = 0; i < n_records; ++i) p += rec_len[i] + 1;
    return p;
}

int tpc_session_create(const tpc_params* params, void* stream, tpc_session** out) {
    if (!params || !out) return set_error("null argument");
    if (params->k == 0 || params->k % 2 == 0) return set_error("value of K must be odd");
    if (params->k > TPC_MAX_K)
        return set_error("The value of K is too big. Please refer to documentaion how to increase the max supported value of K.");
    if (params->filter_bits > 40 || params->filter_bits == 0) return set_error("filter size must be in 1..40 bits");
    if (params->rounds == 0 || params->shard_count == 0 || params->shard_index >= params->shard_count)
        return set_error("bad rounds / shard parameters");
    if (params->q == 0) return set_error("the number of hash functions must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_error("no CUDA device: twopaco_b200 has no CPU fallback (%s)", cudaGetErrorString(e));
    tpc_session* s = new (std::nothrow) tpc_session();
    if (!s) return set_error("out of memory");
    s->prm = *params;
    s->stream = (cudaStream_t)stream;
    s->W = (int)((params->k + 31) / 32);
    s->filter_bits_eff = std::max<uint32_t>(params->filter_bits, 9u);
    if (const char* e = getenv("TPC_FILTER_MODE")) s->filter_mode = !strcmp(e, "direct") ? 1 : !strcmp(e, "binned") ? 2 : 0;
    if (const char* e = getenv("TPC_INLINE_KEYS")) s->allow_inline = atoi(e) != 0;
    if (const char* e = getenv("TPC_SLICE_LOG2")) s->slice_log2 = std::min(31, std::max(8, atoi(e)));
    if (const char* e = getenv("TPC_BIN_BUFFER_MB")) s->bin_budget_bytes = (uint64_t)atoll(e) << 20;
    if (const char* e = getenv("TPC_SUBROUNDS")) s->sub_rounds_env = std::min(64, std::max(0, atoi(e)));
    if (const char* e = getenv("TPC_PIPELINE")) s->pipe_env = atoi(e);
    if (const char* e = getenv("TPC_PIPE_BIN_CTAS")) s->pipe_bin_ctas = std::max(1, atoi(e));
    if (const char* e = getenv("TPC_PIPE_FILL_CTAS")) s->pipe_fill_ctas = std::max(1, atoi(e));
    if (const char* e = getenv("TPC_BIN_CTAS")) s->bin_ctas = atoi(e);
    if (const char* e = getenv("TPC_APPLY_CTAS")) s->apply_ctas = atoi(e);
    s->rounds_eff = params->rounds;
    cudaGetDevice(&s->device);
    configure_pool(s->device);
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    for (auto& ev : s->ev) cudaEventCreate(&ev);
    if (dev_alloc(&s->d_ctr, sizeof(Counters), s->stream) != cudaSuccess || dev_alloc(&s->d_id, sizeof(long long), s->stream) != cudaSuccess ||
        dev_alloc(&s->d_hll, (4u << kHllBits), s->stream) != cudaSuccess) {
        tpc_session_destroy(s);
        return set_error("cudaMalloc failed");
    }
    cudaMemsetAsync(s->d_ctr, 0, sizeof(Counters), s->stream);
    *out = s;
    return 0;
}

void tpc_session_destroy(tpc_session* s) {
    if (!s) return;
    trace("session destroy: enter");
    cudaStreamSynchronize(s->stream);
    if (s->windowed) { s->d_mask = nullptr; s->d_stubmask = nullptr; }   // (views of d_wmask / d_wstub)
    void* ptrs[] = {s->d_codes, s->d_nmask, s->d_rec_start, s->d_rec_len, s->d_sep_before, s->d_filter, s->d_mask,
                    s->d_stubmask, s->d_T, s->d_J, s->d_local, s->d_sorted, s->d_sort_tmp, s->d_ctr, s->d_id,
                    s->d_tile_rec, s->d_tile_stub, s->d_scan_scratch, s->d_marklist, s->d_marklist_counts, s->d_skew, s->d_wmask, s->d_wstub, s->d_local_keys, s->d_bin_rec, s->d_bin_count, s->d_bin_ov, s->d_hll, s->d_own_extra};
    for (void* p : ptrs)
        dev_free(p, s->stream);
    cudaStreamSynchronize(s->stream);
    for (auto& ev : s->ev)
        if (ev) cudaEventDestroy(ev);
    if (s->up_ev_owned)
        for (auto& ev : s->up_ev)
            if (ev) cudaEventDestroy(ev);
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
    if (s->bin_stream) { cudaStreamSynchronize(s->bin_stream); cudaStreamDestroy(s->bin_stream); }
    for (auto& ev : s->pipe_ev)
        if (ev) cudaEventDestroy(ev);
    delete s;
    trace("session destroy: done");
}

static int adopt_records(tpc_session* s, const tpc_genome* g) {
    if (g->n_positions >= (1ull << kPosBits)) return set_error("genome too large: %llu positions", (unsigned long long)g->n_positions);
    s->rec_start.assign(g->rec_start, g->rec_start + g->n_records);
    s->rec_len.assign(g->rec_len, g->rec_len + g->n_records);
    s->sep_before.assign(g->n_records, 0);
    s->emit_prev.assign(g->n_records + 1, 0);
    uint64_t prev = 0;  // JunctionPositionWriter::nowChr_ (junctionapi.h:109,120-123)
    for (uint64_t i = 0; i < g->n_records; ++i) {
        if (g->rec_len[i] >> 32) return set_error("record %llu is longer than 2^32 bp", (unsigned long long)i);
        s->emit_prev[i] = prev;
        if (g->rec_len[i] >= s->prm.k) {
            s->sep_before[i] = (uint32_t)(i - prev);
            prev = i;
        }
    }
    s->emit_prev[g->n_records] = prev;
    s->ntiles = (g->n_positions + kTilePos - 1) / kTilePos;
    size_t nr = std::max<uint64_t>(g->n_records, 1);
    CK(dev_alloc(&s->d_rec_start, nr * 8, s->stream));
    CK(dev_alloc(&s->d_rec_len, nr * 8, s->stream));
    CK(dev_alloc(&s->d_sep_before, nr * 4, s->stream));
    if (g->n_records) {
        CK(cudaMemcpyAsync(s->d_rec_start, s->rec_start.data(), g->n_records * 8, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_rec_len, s->rec_len.data(), g->n_records * 8, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_sep_before, s->sep_before.data(), g->n_records * 4, cudaMemcpyHostToDevice, s->stream));
    }
    return 0;
}

// make the compute stream wait for the upload chunks that cover tiles [0, tile_end] (the kernels
// read a few words past the end of a tile)
static int wait_genome(tpc_session* s, uint64_t tile_end) {
    while (s->up_waited < s->up_ev.size() && s->up_tile_begin[s->up_waited] <= tile_end) {
        CK(cudaStreamWaitEvent(s->stream, s->up_ev[s->up_waited], 0));
        ++s->up_waited;
    }
    return 0;
}

// The n-mask (1 bit per position) is a third of the packed genome's bytes and almost everywhere uniform: zero inside the
// sequences, ones only at record separators, N runs and the padding.  Classify it in blocks of 64 KiB (all host threads,
// one upload chunk at a time, while the previous chunk's words cross PCIe): runs of all-zero / all-one blocks become a
// cudaMemsetAsync on the device, only the mixed blocks are copied.  -> block kinds of words [m0, m1): 0 / 1 = uniform, 2 = mixed
constexpr uint64_t kMaskBlockWords = 8192;
static void classify_mask_blocks(const uint64_t* n_mask, uint64_t m0, uint64_t m1, std::vector<uint8_t>* kind) {
    const uint64_t nb = (m1 - m0 + kMaskBlockWords - 1) / kMaskBlockWords;
    kind->assign(nb, 2);
    auto work = [&](uint64_t b0, uint64_t b1) {
        for (uint64_t b = b0; b < b1; ++b) {
            const uint64_t lo = m0 + b * kMaskBlockWords, hi = std::min(m1, lo + kMaskBlockWords);
            uint64_t any = 0, all = ~0ull;
            for (uint64_t i = lo; i < hi; ++i) { any |= n_mask[i]; all &= n_mask[i]; }
            (*kind)[b] = any == 0 ? 0 : all == ~0ull ? 1 : 2;
        }
    };
    static const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    const unsigned nt = (unsigned)std::min<uint64_t>(hw, std::max<uint64_t>(1, nb / 64));
    if (nt <= 1) { work(0, nb); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, nb * t / nt, nb * (t + 1) / nt);
    for (std::thread& t : pool) t.join();
}

int tpc_session_set_genome_host(tpc_session* s, const tpc_genome* g) {
    if (!s || !g) return set_error("null argument");
    if (s->g.codes) return set_error("genome already set");
    uint64_t cw = tpc_code_words(g->n_positions), mw = tpc_mask_words(g->n_positions);
    CK(dev_alloc(&s->d_codes, cw * 8, s->stream));
    CK(dev_alloc(&s->d_nmask, mw * 8, s->stream));
    s->g = GenomeView{s->d_codes, s->d_nmask, g->n_positions};
    if (int rc = adopt_records(s, g)) return rc;
    // Upload in position order on a second stream, one event per chunk: the first pass over the
    // genome (k_bin / k_own, wave by wave) starts as soon as the chunks it reads have arrived.
    if (!s->copy_stream) CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventRecord(s->ev[9], s->stream));                 // allocations are stream-ordered
    CK(cudaStreamWaitEvent(s->copy_stream, s->ev[9], 0));
    const uint64_t chunks = std::max<uint64_t>(1, std::min<uint64_t>(16, s->ntiles / 512));
    const bool sparse_mask = !(getenv("TPC_SPARSE_MASK") && atoi(getenv("TPC_SPARSE_MASK")) == 0);   // (=0: copy everything, tests / tuning)
    std::vector<uint8_t> kind;
    s->h2d_bytes = 0;
    for (uint64_t c = 0; c < chunks; ++c) {
        uint64_t t0 = s->ntiles * c / chunks, t1 = s->ntiles * (c + 1) / chunks;
        uint64_t c0 = t0 * kTileThreads, c1 = c + 1 == chunks ? cw : t1 * kTileThreads;
        uint64_t m0 = t0 * (kTileThreads / 2), m1 = c + 1 == chunks ? mw : t1 * (kTileThreads / 2);
        CK(cudaMemcpyAsync(s->d_codes + c0, g->codes + c0, (c1 - c0) * 8, cudaMemcpyHostToDevice, s->copy_stream));
        s->h2d_bytes += (c1 - c0) * 8;
        if (!sparse_mask) {
            CK(cudaMemcpyAsync(s->d_nmask + m0, g->n_mask + m0, (m1 - m0) * 8, cudaMemcpyHostToDevice, s->copy_stream));
            s->h2d_bytes += (m1 - m0) * 8;
        } else {
            classify_mask_blocks(g->n_mask, m0, m1, &kind);   // (the codes of this chunk are crossing PCIe meanwhile)
            for (uint64_t b = 0; b < kind.size();) {
                uint64_t e = b + 1;
                while (e < kind.size() && kind[e] == kind[b]) ++e;
                const uint64_t lo = m0 + b * kMaskBlockWords, hi = std::min(m1, m0 + e * kMaskBlockWords);
                if (kind[b] == 2) {
                    CK(cudaMemcpyAsync(s->d_nmask + lo, g->n_mask + lo, (hi - lo) * 8, cudaMemcpyHostToDevice, s->copy_stream));
                    s->h2d_bytes += (hi - lo) * 8;
                } else CK(cudaMemsetAsync(s->d_nmask + lo, kind[b] ? 0xFF : 0, (hi - lo) * 8, s->copy_stream));
                b = e;
            }
        }
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaEventRecord(e, s->copy_stream));
        s->up_ev.push_back(e);
        s->up_tile_begin.push_back(t0);
    }
    s->up_waited = 0;
    return 0;
}

int tpc_session_set_genome_device(tpc_session* s, const tpc_genome* g) {
    if (!s || !g) return set_error("null argument");
    if (s->g.codes) return set_error("genome already set");
    s->g = GenomeView{g->codes, g->n_mask, g->n_positions};
    return adopt_records(s, g);
}

int tpc_session_add_genome_event(tpc_session* s, uint64_t tile_begin, void* cuda_event) {
    if (!s || !s->g.codes || !cuda_event) return set_error("set_genome_device first");
    if (s->d_codes) return set_error("the genome was uploaded by the session itself");
    if (s->have_candidates || s->up_waited) return set_error("genome events must be added before find_candidates");
    if (s->up_ev.empty() ? tile_begin != 0 : tile_begin <= s->up_tile_begin.back())
        return set_error("genome events must be added in ascending tile order, starting at tile 0");
    s->up_ev_owned = false;
    s->up_ev.push_back((cudaEvent_t)cuda_event);
    s->up_tile_begin.push_back(tile_begin);
    return 0;
}

static double hll_estimate(const std::vector<uint32_t>& reg) {
    const double m = (double)reg.size();
    double sum = 0;
    uint32_t zeros = 0;
    for (uint32_t r : reg) { sum += std::ldexp(1.0, -(int)r); zeros += r == 0; }
    double e = (0.7213 / (1.0 + 1.079 / m)) * m * m / sum;
    if (e <= 2.5 * m && zeros) e = m * std::log(m / zeros);  // small-range correction (linear counting)
    return e;
}

// the binned path needs 2..256 filter slices and an input large enough to be worth it
static bool binned_applies(const tpc_session* s, BinView* out) {
    if (s->filter_mode == 1) return false;
    int bb = (int)s->filter_bits_eff - 3 - (int)s->slice_log2;
    if (bb < 1 || bb > 8) return false;
    if (s->filter_mode == 0 && s->g.npos < (1ull << 22)) return false;
    uint32_t sib = s->filter_bits_eff - 8 - (uint32_t)bb;
    if (sib > (uint32_t)kBinCodeShift) return false;
    if (out) { out->bucket_bits = (uint32_t)bb; out->sib_bits = sib; }
    return true;
}

static uint32_t own_planes_for(uint32_t rounds_local) {  // bits needed for ids 1..rounds_local; 0 = no shared scan
    if (rounds_local > 15) return 0;
    uint32_t p = 1;
    while ((1u << p) <= rounds_local) ++p;
    return p;
}

// Sub-rounds: a round whose records (12 B per owned k-mer) fit free HBM in ONE wave is binned once
// for both filter passes, and its filter holds 1/S of the vertices, so Bloom false positives (the
// marks pass 2 has to remove) drop steeply.  -r and -f keep their meaning; like -r, the split is
// unobservable in the output.
static uint32_t choose_sub_rounds(tpc_session* s) {
    s->pipe = false;
    const bool can_pipe = s->pipe_env && binned_applies(s, nullptr) && !s->bin_budget_bytes;
    const uint64_t avail = available_bytes(s->device) + s->kept_scratch_bytes() + s->own_extra_bytes;   // (what a first call sees)
    const uint64_t plane_bytes = s->ntiles * kTileThreads * 4;
    const uint64_t base = (uint64_t)s->prm.rounds * s->prm.shard_count;
    // does the record scratch of one round (x2 when two rounds are in flight) fit with S sub-rounds?
    auto fits = [&](uint32_t S, int scratches) {
        uint32_t planes = own_planes_for(s->prm.rounds * S);
        uint64_t extra = planes > 1 ? (planes - 1) * plane_bytes : 0;
        if (avail <= extra) return false;
        return (double)(s->g.npos / (base * S)) * 14.0 * scratches <= (double)(avail - extra) * 0.88;  // (binning takes 92 %)
    };
    if (s->sub_rounds_env) {
        const uint32_t S = (uint32_t)s->sub_rounds_env;
        s->pipe = can_pipe && s->prm.rounds * S >= 2 && own_planes_for(s->prm.rounds * S) > 0 && fits(S, 2);
        return S;
    }
    if (s->bin_budget_bytes || !binned_applies(s, nullptr)) return 1;
    uint32_t S1 = 8;
    for (uint32_t S = 1; S <= 8; ++S)
        if (fits(S, 1)) { S1 = S; break; }
    // k_bin_list stages the owned positions of an 8192-position tile 2048 (or 1024) at a time: a share of 1/3 would run
    // a full and a one-third-full staging round per tile (measured at C3 on one GPU: 3 sub-rounds bin in 241 ms, 4 in
    // 218 ms), so the split is rounded up to a power of two -- whole stages for every tile, whatever the input
    while (S1 & (S1 - 1)) ++S1;
    if (can_pipe && s->prm.rounds * S1 >= 2) {
        // Pipelined rounds (opt-in, TPC_PIPELINE=1): two half-size scratches, round r+1 binned beside the fill of round r.
        // Measured at C3 on one GPU the overlap gains nothing any more (the fill is bound by L2 atomic throughput and loses
        // as much as the binning hides: 590 ms pipelined with 5 sub-rounds, 580 ms in sequence with 4), hence not the default.
        for (uint32_t S = S1; S <= 2 * S1 && S <= 12; ++S)
            if (s->prm.rounds * S >= 2 && own_planes_for(s->prm.rounds * S) > 0 && fits(S, 2)) {
                s->pipe = true;
                return S;
            }
    }
    return S1;
}

// first tile boundary in (from, to] at which another upload chunk has to be waited for
static uint64_t next_cut(const tpc_session* s, uint64_t from, uint64_t to) {
    for (size_t c = s->up_waited; c < s->up_ev.size(); ++c)
        if (s->up_tile_begin[c] > from) return std::min(to, s->up_tile_begin[c]);
    return to;
}

// Record scratch of the binned path: sized once per find_candidates call (every round owns the same
// share of the positions) and kept until the call ends, so the rounds do not go through the allocator.
// Returns -1 when the binned path does not apply, 0 on success, >0 on error.
static int binned_setup(tpc_session* s, const KParams& kp) {
    if (s->bin_ready) return 0;
    BinView bv{};
    if (!binned_applies(s, &bv)) return -1;
    bv.q = kp.q;
    const uint32_t buckets = 1u << bv.bucket_bits;
    uint64_t budget = s->bin_budget_bytes ? s->bin_budget_bytes : (uint64_t)((available_bytes(s->device) + s->kept_scratch_bytes()) * 0.92);
    uint64_t wave_tiles = 0, nwaves = 0;
    for (int attempt = 0;; ++attempt) {
        // records of one wave must fit the budget: 12 B per record + 8 % slack per slice + overflow area
        uint64_t max_records = budget / 14;
        // positions inside a wave are 32 + (25 - sib_bits) bits wide (tpc_bin.cuh record layout)
        uint64_t wave_pos = std::min<uint64_t>(s->ntiles * (uint64_t)kTilePos, (1ull << (32 + kBinCodeShift - bv.sib_bits)) - kTilePos);
        if (max_records < wave_pos / kp.nparts) wave_pos = std::max<uint64_t>(max_records * kp.nparts, kTilePos);
        wave_tiles = std::max<uint64_t>(wave_pos / kTilePos, 1);
        nwaves = (s->ntiles + wave_tiles - 1) / wave_tiles;
        wave_tiles = (s->ntiles + nwaves - 1) / nwaves;  // equal waves
        uint64_t est = wave_tiles * kTilePos / kp.nparts;
        bv.cap = ((uint64_t)(est / buckets * 1.08) + 8192 + 31) / 32 * 32;
        bv.ov_cap = std::max<uint64_t>(1 << 16, est / 64);
        if (s->pipe && nwaves > 1) s->pipe = false;   // (cannot happen with the sub-rounds chosen for it, but stay safe)
        const uint32_t halves = s->pipe ? 2 : 1;
        const uint64_t need = (uint64_t)halves * buckets * 3 * bv.cap * 4;
        if (s->d_bin_rec && s->bin_rec_bytes == need) break;   // kept from the previous call
        if (s->d_bin_rec) { CK(dev_free(s->d_bin_rec, s->stream)); s->d_bin_rec = nullptr; }
        s->bin_rec_bytes = need;
        cudaError_t e = dev_alloc(&s->d_bin_rec, s->bin_rec_bytes, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e == cudaErrorMemoryAllocation && attempt < 4 && !s->bin_budget_bytes) {
            cudaGetLastError();  // not enough contiguous memory after all: one scratch, then smaller waves
            invalidate_memory_budget();
            s->d_bin_rec = nullptr;
            s->bin_rec_bytes = 0;
            if (s->pipe) s->pipe = false;
            else budget = budget * 6 / 10;
            continue;
        }
        CK(e);
        break;
    }
    const uint32_t halves = s->pipe ? 2 : 1;
    const uint64_t count_bytes = (uint64_t)halves * (buckets + 1) * 8, ov_bytes = (uint64_t)halves * bv.ov_cap * 16;
    if (!s->d_bin_count || s->bin_count_bytes != count_bytes) {
        if (s->d_bin_count) CK(dev_free(s->d_bin_count, s->stream));
        CK(dev_alloc(&s->d_bin_count, count_bytes, s->stream));
        s->bin_count_bytes = count_bytes;
    }
    if (!s->d_bin_ov || s->bin_ov_bytes != ov_bytes) {
        if (s->d_bin_ov) CK(dev_free(s->d_bin_ov, s->stream));
        CK(dev_alloc(&s->d_bin_ov, ov_bytes, s->stream));
        s->bin_ov_bytes = ov_bytes;
    }
    {   // keep the scratch for the next call when at least a quarter of the device stays free beside it (index, image, tables)
        uint64_t total_b = 0;
        const uint64_t avail = available_bytes(s->device, &total_b);
        s->bin_keep = !s->bin_budget_bytes && avail >= total_b / 4 && !(getenv("TPC_KEEP_SCRATCH") && atoi(getenv("TPC_KEEP_SCRATCH")) == 0);
    }
    bv.rec = s->d_bin_rec; bv.count = s->d_bin_count; bv.ov_count = s->d_bin_count + buckets; bv.ov = s->d_bin_ov;
    s->bin_view = bv;
    for (uint32_t h = 0; h < 2; ++h) {
        BinView hv = bv;
        if (h < halves) {
            hv.rec = s->d_bin_rec + (uint64_t)h * buckets * 3 * bv.cap;
            hv.count = s->d_bin_count + (uint64_t)h * (buckets + 1);
            hv.ov_count = hv.count + buckets;
            hv.ov = s->d_bin_ov + (uint64_t)h * bv.ov_cap * 4;
        }
        s->pipe_view[h] = hv;
        s->pipe_round[h] = -1;
    }
    if (s->pipe) {
        if (!s->bin_stream) CK(cudaStreamCreateWithFlags(&s->bin_stream, cudaStreamNonBlocking));
        for (auto& ev : s->pipe_ev)
            if (!ev) CK(cudaEventCreate(&ev));
    }
    s->bin_wave_tiles = wave_tiles;
    s->bin_nwaves = nwaves;
    s->bin_ready = true;
    return 0;
}

static int binned_release(tpc_session* s) {
    s->bin_ready = false;
    if (s->bin_keep && s->d_bin_rec) return 0;   // (binned_setup of the next call reuses or replaces it; the session's destructor frees it)
    for (void** p : {(void**)&s->d_bin_rec, (void**)&s->d_bin_count, (void**)&s->d_bin_ov}) {
        if (*p) CK(dev_free(*p, s->stream));
        *p = nullptr;
    }
    s->bin_rec_bytes = 0; s->bin_count_bytes = 0; s->bin_ov_bytes = 0;
    return 0;
}

// Filter passes of one round through the binned path.  Returns -1 when the binned path does not
// apply and -2 when a slice overflowed beyond the overflow area (then the caller uses k_fill /
// k_query), 0 on success, >0 on error.
static int filter_passes_binned(tpc_session* s, const KParams& kp, float* ms_bin, float* ms_fill, float* ms_query) {
    trace("filter passes (binned): enter");
    if (int rc = binned_setup(s, kp)) return rc;
    trace("record scratch ready");
    LaunchCtx lc = s->lctx();
    BinView bv = s->bin_view;       // (uniform layout; a skewed round switches this copy to per-slice capacities)
    SliceLayout sl;
    bool skew_layout = false;
    const uint32_t buckets = 1u << bv.bucket_bits;
    const uint64_t wave_tiles = s->bin_wave_tiles, nwaves = s->bin_nwaves;
    Events evs(3);
    cudaEvent_t e0 = evs[0], e1 = evs[1], e2 = evs[2];
    unsigned long long ov_total = 0, ov_now = 0;
    auto finish_wave = [&](float* a, float* b) -> int {
        CK(cudaMemcpyAsync(&ov_now, bv.ov_count, 8, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaEventSynchronize(e2));
        CK(cudaStreamSynchronize(s->stream));
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1); *a += t;
        cudaEventElapsedTime(&t, e1, e2); *b += t;
        ov_total = std::max(ov_total, ov_now);
        return 0;
    };
    // bin the tiles [t0, t1): cut at the upload chunks still in flight so that the first pass over a
    // host genome overlaps its upload; sharded rounds read the ownership planes (one k_own scan for all
    // the rounds of this call when they fit kMaxOwnPlanes, else one scan per round)
    const bool sharded = kp.nparts > 1;
    const uint32_t part_base = s->prm.shard_index * s->rounds_eff;
    OwnPlanes op = s->own;
    op.id = s->own_shared ? kp.part - part_base + 1 : 1;
    if (!s->own_shared) op.n = 1;
    // a GPU that runs one round in one wave reads its ownership exactly once: k_bin_list computes it itself (no k_own, no plane)
    const bool fused_own = sharded && s->rounds_eff == 1 && nwaves == 1 && !(getenv("TPC_FUSED_OWN") && atoi(getenv("TPC_FUSED_OWN")) == 0);
    if (fused_own) op.n = 0;
    auto bin_range = [&](uint64_t t0, uint64_t t1, uint64_t base) -> int {
        for (uint64_t a = t0; a < t1;) {
            const uint64_t b = next_cut(s, a, t1);
            if (int wrc = wait_genome(s, b)) return wrc;
            if (sharded && !s->own_shared && !fused_own) CK(W_DISPATCH(s, own(lc, s->g, kp, kp.part, 1, a, b, op)));
            if (sharded && s->own_shared && !fused_own && s->own_done_tiles < b) {
                CK(W_DISPATCH(s, own(lc, s->g, kp, part_base, s->rounds_eff, s->own_done_tiles, b, s->own)));
                s->own_done_tiles = b;
            }
            CK(W_DISPATCH(s, bin(lc, s->g, kp, bv, a, b, base, sharded ? &op : nullptr)));
            a = b;
        }
        return 0;
    };
    int rc = 0;
    for (int pass = 0; pass < 2 && rc == 0; ++pass) {           // 0 = fill, 1 = query
        for (uint64_t wv = 0; wv < nwaves && rc == 0; ++wv) {
            uint64_t t0 = wv * wave_tiles, t1 = std::min(s->ntiles, t0 + wave_tiles);
            uint64_t base = t0 * kTilePos;
            bool rebin = !(pass == 1 && nwaves == 1);           // one wave: the records serve both passes
            CK(cudaEventRecord(e0, s->stream));
            if (rebin) {
                CK(cudaMemsetAsync(s->d_bin_count, 0, (buckets + 1) * 8, s->stream));
                if (int brc = bin_range(t0, t1, base)) return brc;
            }
            CK(cudaEventRecord(e1, s->stream));
            trace(rebin ? "binning enqueued" : "(records shared with the fill pass)");
            for (uint32_t b = 0; b < buckets; ++b) {
                if (pass == 0) CK(launch_apply_fill(lc, s->d_filter, bv, sl, b, s->d_ctr));
                else CK(launch_apply_query(lc, s->d_filter, bv, sl, b, s->d_mask, base, s->d_ctr, s->d_hll, s->mark_list(true)));
            }
            CK(launch_apply_overflow(lc, s->d_filter, bv, pass, s->d_mask, base, s->d_ctr, s->d_hll, s->mark_list(pass == 1)));
            CK(cudaEventRecord(e2, s->stream));
            trace("apply kernels enqueued");
            rc = finish_wave(ms_bin, pass == 0 ? ms_fill : ms_query);
            trace(pass == 0 ? "fill pass synchronised" : "query pass synchronised");
            if (rc == 0 && pass == 0 && nwaves == 1 && ov_now > bv.ov_cap && !skew_layout) {
                // Skewed input: slices ran over their arrays AND the overflow list.  The reservation counters hold the exact
                // number of records of every slice: re-bin this round into arrays of exactly those sizes (one more binning
                // pass instead of the direct kernels for the whole round) and fill again -- the bits already set are correct.
                std::vector<unsigned long long> cnt(buckets);
                CK(cudaMemcpyAsync(cnt.data(), bv.count, buckets * 8, cudaMemcpyDeviceToHost, s->stream));
                CK(cudaStreamSynchronize(s->stream));
                s->skew_off.assign(buckets, 0);
                s->skew_cap.assign(buckets, 0);
                unsigned long long total = 0;
                for (uint32_t b = 0; b < buckets; ++b) {
                    s->skew_cap[b] = (cnt[b] + cnt[b] / 64 + 1024 + 31) / 32 * 32;
                    s->skew_off[b] = total;
                    total += s->skew_cap[b];
                }
                if (total * 12 <= s->bin_rec_bytes) {
                    if (!s->d_skew) CK(dev_alloc(&s->d_skew, (uint64_t)2 * kBinMaxBuckets * 8, s->stream));
                    CK(cudaMemcpyAsync(s->d_skew, s->skew_off.data(), buckets * 8, cudaMemcpyHostToDevice, s->stream));
                    CK(cudaMemcpyAsync(s->d_skew + kBinMaxBuckets, s->skew_cap.data(), buckets * 8, cudaMemcpyHostToDevice, s->stream));
                    bv.off = s->d_skew;
                    bv.capv = s->d_skew + kBinMaxBuckets;
                    sl.off = &s->skew_off;
                    sl.capv = &s->skew_cap;
                    skew_layout = true;
                    ++s->skew_rebins;
                    ov_total = 0;
                    --wv;            // the same wave again, with the new layout
                    continue;
                }
            }
        }
    }
    if (rc) return rc;
    if (ov_total > bv.ov_cap) return -2;  // heavily skewed input: the caller redoes the round with k_fill / k_query
    s->used_binned = true;
    s->st.bin_waves = (uint32_t)nwaves;
    return 0;
}

// Bin round `part` (ownership id `id`) into the half scratch `bv` on stream `lc.stream`.
static int bin_whole_round(tpc_session* s, const LaunchCtx& lc, const KParams& kp, const BinView& bv) {
    const uint32_t buckets = 1u << bv.bucket_bits;
    const uint32_t part_base = s->prm.shard_index * s->rounds_eff;
    OwnPlanes op = s->own;
    op.id = kp.part - part_base + 1;
    CK(cudaMemsetAsync(bv.count, 0, (buckets + 1) * 8, lc.stream));
    for (uint64_t a = 0; a < s->ntiles;) {   // (cuts only while the genome is still being uploaded: first round, main stream)
        const uint64_t b = lc.stream == s->stream ? next_cut(s, a, s->ntiles) : s->ntiles;
        if (lc.stream == s->stream)
            if (int wrc = wait_genome(s, b)) return wrc;
        if (s->own_done_tiles < b) {
            CK(W_DISPATCH(s, own(lc, s->g, kp, part_base, s->rounds_eff, s->own_done_tiles, b, s->own)));
            s->own_done_tiles = b;
        }
        CK(W_DISPATCH(s, bin(lc, s->g, kp, bv, a, b, 0, &op)));
        a = b;
    }
    return 0;
}

// Filter passes of round r (of rounds_eff) with the next round's binning overlapped (see tpc_session::pipe).
// Same return convention as filter_passes_binned.
static int filter_passes_pipelined(tpc_session* s, uint32_t r, const KParams& kp, float* ms_bin, float* ms_fill, float* ms_query) {
    if (int rc = binned_setup(s, kp)) return rc;
    if (!s->pipe) return -3;   // the scratch could not be split after all: caller uses the one-scratch path
    const int h = (int)(r & 1);
    const BinView bv = s->pipe_view[h];
    const uint32_t buckets = 1u << bv.bucket_bits;
    const bool has_next = r + 1 < s->rounds_eff;
    LaunchCtx lc = s->lctx();
    Events evs(4);
    cudaEvent_t e0 = evs[0], e1 = evs[1], e2 = evs[2], e3 = evs[3];
    float t_bin_ahead = 0;
    CK(cudaEventRecord(e0, s->stream));
    if (s->pipe_round[h] != (long long)r) {   // first round of the call (or after a fallback): bin here, nothing to overlap with
        if (int rc = bin_whole_round(s, lc, kp, bv)) return rc;
        s->pipe_round[h] = r;
    } else {
        CK(cudaStreamWaitEvent(s->stream, s->pipe_ev[h], 0));
    }
    CK(cudaEventRecord(e1, s->stream));
    if (has_next) {
        // everything enqueued on the main stream so far (ownership planes, the genome upload, the previous round's
        // readers of the other half) precedes the next round's binning
        if (int wrc = wait_genome(s, s->ntiles)) return wrc;
        CK(cudaEventRecord(s->pipe_ev[2], s->stream));
        CK(cudaStreamWaitEvent(s->bin_stream, s->pipe_ev[2], 0));
        LaunchCtx lb{s->bin_stream, s->sm_count, &s->launches, s->pipe_bin_ctas, 0};
        KParams kn = s->kparams(kp.part + 1);
        CK(cudaEventRecord(s->pipe_ev[3], s->bin_stream));
        if (int rc = bin_whole_round(s, lb, kn, s->pipe_view[h ^ 1])) return rc;
        CK(cudaEventRecord(s->pipe_ev[4], s->bin_stream));
        CK(cudaEventRecord(s->pipe_ev[h ^ 1], s->bin_stream));
        s->pipe_round[h ^ 1] = r + 1;
        lc.apply_ctas = s->pipe_fill_ctas;   // share the SMs with the binning kernel
    }
    const SliceLayout sl;   // (pipelined rounds keep the uniform layout; a skewed round falls back)
    for (uint32_t b = 0; b < buckets; ++b) CK(launch_apply_fill(lc, s->d_filter, bv, sl, b, s->d_ctr));
    CK(launch_apply_overflow(lc, s->d_filter, bv, 0, s->d_mask, 0, s->d_ctr, s->d_hll, s->mark_list(false)));
    CK(cudaEventRecord(e2, s->stream));
    if (has_next) CK(cudaStreamWaitEvent(s->stream, s->pipe_ev[h ^ 1], 0));   // the query runs alone, at full occupancy
    lc.apply_ctas = s->apply_ctas;
    for (uint32_t b = 0; b < buckets; ++b) CK(launch_apply_query(lc, s->d_filter, bv, sl, b, s->d_mask, 0, s->d_ctr, s->d_hll, s->mark_list(true)));
    CK(launch_apply_overflow(lc, s->d_filter, bv, 1, s->d_mask, 0, s->d_ctr, s->d_hll, s->mark_list(true)));
    CK(cudaEventRecord(e3, s->stream));
    unsigned long long ov_now = 0;
    CK(cudaMemcpyAsync(&ov_now, bv.ov_count, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1); *ms_bin += t;       // binning done here (first round) or the wait for the overlapped one
    cudaEventElapsedTime(&t, e1, e2); *ms_fill += t;      // (overlapped with the next round's binning)
    cudaEventElapsedTime(&t, e2, e3); *ms_query += t;     // (includes waiting for that binning to end)
    if (has_next && cudaEventElapsedTime(&t_bin_ahead, s->pipe_ev[3], s->pipe_ev[4]) == cudaSuccess) s->st.ms_bin_overlapped += t_bin_ahead;
    if (ov_now > bv.ov_cap) return -2;
    s->used_binned = true;
    s->st.bin_waves = 1;
    return 0;
}

static uint64_t padded_mask_words(const tpc_session* s) {
    const uint64_t n = std::max<uint32_t>(s->prm.shard_count, 1);
    const uint64_t chunk_tiles = (s->ntiles + n - 1) / n;
    return std::max<uint64_t>(chunk_tiles * n * kTileThreads, 1);
}

namespace tpc { static int find_candidates_windowed(tpc_session* s); }

int tpc_session_find_candidates(tpc_session* s) {
    if (s && s->windowed) return tpc::find_candidates_windowed(s);
    if (!s || !s->g.codes) return set_error("no genome set");
    trace("find_candidates: enter");
    LaunchCtx lc = s->lctx();
    const uint64_t mask_words = s->ntiles * kTileThreads;
    const uint64_t filter_bytes = (1ull << s->filter_bits_eff) / 8;
    if (!s->d_filter) CK(dev_alloc(&s->d_filter, filter_bytes, s->stream));
    // the candidate mask is padded to shard_count equal tile-aligned chunks so that the shards' masks can be
    // reduce-scattered straight into the position slices the GPUs emit (tpc_session_candidate_mask)
    const uint64_t mask_words_padded = padded_mask_words(s);
    if (!s->d_mask) {
        CK(dev_alloc(&s->d_mask, mask_words_padded * 4, s->stream));
        CK(dev_alloc(&s->d_stubmask, std::max<uint64_t>(mask_words, 1) * 4, s->stream));
    }
    CK(cudaMemsetAsync(s->d_mask, 0, mask_words_padded * 4, s->stream));
    CK(cudaMemsetAsync(s->d_ctr, 0, sizeof(Counters), s->stream));
    s->local_count = 0;
    s->st = tpc_stats{};
    s->st.positions = s->g.npos;
    WallTimer wall(&s->st.ms_wall_candidates);
    const bool verbose = getenv("TPC_VERBOSE") != nullptr;
    const double t_start = now_ms();
    auto vlog = [&](const char* what, int round) {
        if (verbose) fprintf(stderr, "[tpc find_candidates] +%9.3f ms  %s (round %d)\n", now_ms() - t_start, what, round);
    };
    float ms_bin = 0, ms_fill = 0, ms_query = 0, ms_insert = 0, ms_classify = 0;
    Counters prev{}, cur{};
    // Mark list: room for npos / 32 / shards candidates (C3: 1.5 % of the positions are marks); a denser input overflows
    // it and the exact pass walks the mask instead.  Reading the k-mer of a listed position is a random access into the
    // packed genome, walking the mask stages whole tiles: measured at C3 the list wins below one mark per ~128 positions
    // (4 and more hash-range shards: 13 -> 4 ms on 8 GPUs) and loses above (one GPU: 22 -> 36 ms), so it is kept for
    // 4 and more shards only (TPC_MARK_LIST=0 / 1 overrides).
    const char* ml_env = getenv("TPC_MARK_LIST");
    const bool want_list = ml_env ? atoi(ml_env) != 0 : s->prm.shard_count >= 4;
    if (!want_list && s->d_marklist) {
        CK(dev_free(s->d_marklist, s->stream));
        CK(dev_free(s->d_marklist_counts, s->stream));
        s->d_marklist = nullptr; s->d_marklist_counts = nullptr; s->marklist_regions = 0;
    }
    if (binned_applies(s, nullptr) && want_list) {
        const uint32_t regions = (uint32_t)s->sm_count * 4;
        uint64_t cap = s->g.npos / 32 / s->prm.shard_count + (1u << 20);
        if (const char* e = getenv("TPC_MARK_LIST_CAP")) cap = (uint64_t)atoll(e);   // (tests: a list that overflows)
        const uint32_t region_cap = (uint32_t)std::min<uint64_t>((cap + regions - 1) / regions, 0x7FFFFFFFu);
        if (!s->d_marklist || regions != s->marklist_regions || region_cap != s->marklist_region_cap) {
            if (s->d_marklist) CK(dev_free(s->d_marklist, s->stream));
            if (s->d_marklist_counts) CK(dev_free(s->d_marklist_counts, s->stream));
            CK(dev_alloc(&s->d_marklist, (uint64_t)regions * region_cap * 8, s->stream));
            CK(dev_alloc(&s->d_marklist_counts, regions * 4, s->stream));
            s->marklist_regions = regions; s->marklist_region_cap = region_cap;
        }
    }
    trace("filter / masks / mark list allocated (enqueued)");
    CK(cudaStreamSynchronize(s->stream));  // the allocations above are visible to available_bytes()
    trace("... and synchronised");
    vlog("buffers allocated, mask cleared", -1);
    s->sub_rounds = choose_sub_rounds(s);
    trace("sub-rounds chosen");
    s->rounds_eff = s->prm.rounds * s->sub_rounds;
    s->st.sub_rounds = s->sub_rounds;
    {   // ownership planes shared by all rounds of this call
        const uint32_t planes = s->rounds_eff * s->prm.shard_count > 1 ? own_planes_for(s->rounds_eff) : 0;
        s->own = OwnPlanes{};
        s->own.p[0] = s->d_stubmask;
        s->own.n = std::max<uint32_t>(planes, 1);
        s->own_shared = planes > 0;
        s->own_done_tiles = 0;
        const uint64_t extra_bytes = planes > 1 && binned_applies(s, nullptr) ? (uint64_t)(planes - 1) * std::max<uint64_t>(mask_words, 1) * 4 : 0;
        if (s->d_own_extra && s->own_extra_bytes != extra_bytes) { CK(dev_free(s->d_own_extra, s->stream)); s->d_own_extra = nullptr; s->own_extra_bytes = 0; }
        if (planes > 1 && binned_applies(s, nullptr)) {
            if (!s->d_own_extra) CK(dev_alloc(&s->d_own_extra, extra_bytes, s->stream));
            s->own_extra_bytes = extra_bytes;
            for (uint32_t j = 1; j < planes; ++j) s->own.p[j] = s->d_own_extra + (uint64_t)(j - 1) * mask_words;
        } else if (planes > 1) {
            s->own_shared = false;  // direct path: ownership is decided inline
            s->own.n = 1;
        }
    }
    // Sub-rounds of ONE user round share the exact pass: their candidates are disjoint hash ranges of the
    // same input, the mask only holds this GPU's marks, so one k_insert scan over the marks of all of them
    // (one table, sized from the HyperLogLog sketch accumulated over the sub-rounds) replaces one scan per
    // sub-round.  With -r > 1 the table stays per round, as in the reference (h:337-338).
    const bool merge_insert = s->prm.rounds == 1 && s->sub_rounds > 1;
    trace("ownership planes ready");
    for (uint32_t r = 0; r < s->rounds_eff; ++r) {
        KParams kp = s->kparams(s->prm.shard_index * s->rounds_eff + r);
        const Counters round_start = cur;
        CK(cudaEventRecord(s->ev[0], s->stream));
        CK(cudaMemsetAsync(s->d_filter, 0, filter_bytes, s->stream));  // h:257: zero-filled each round
        if (!merge_insert || r == 0) {
            CK(cudaMemsetAsync(s->d_hll, 0, 4u << kHllBits, s->stream));
            if (s->d_marklist) CK(cudaMemsetAsync(s->d_marklist_counts, 0, s->marklist_regions * 4, s->stream));
            s->marklist_valid = s->d_marklist != nullptr;
        }
        float b_bin = 0, b_fill = 0, b_query = 0;
        int brc = s->pipe ? filter_passes_pipelined(s, r, kp, &b_bin, &b_fill, &b_query) : -3;
        if (brc == -3) brc = filter_passes_binned(s, kp, &b_bin, &b_fill, &b_query);
        if (brc > 0) return brc;
        if (brc == -2) {  // redo this round from scratch; marks already set are true marks and may stay
            CK(cudaMemsetAsync(s->d_filter, 0, filter_bytes, s->stream));
            CK(cudaMemcpyAsync(s->d_ctr, &round_start, sizeof round_start, cudaMemcpyHostToDevice, s->stream));
            if (!merge_insert) CK(cudaMemsetAsync(s->d_hll, 0, 4u << kHllBits, s->stream));
            CK(cudaEventRecord(s->ev[0], s->stream));
        }
        if (int wrc = wait_genome(s, s->ntiles)) return wrc;
        if (brc < 0) {
            s->marklist_valid = false;   // the direct query kernel marks the mask only
            CK(W_DISPATCH(s, fill(lc, s->g, s->d_filter, kp, 0, s->ntiles, s->d_ctr)));
            CK(cudaEventRecord(s->ev[1], s->stream));
            CK(W_DISPATCH(s, query(lc, s->g, s->d_filter, kp, 0, s->ntiles, s->d_mask, r > 0 || brc == -2, s->d_ctr, s->d_hll)));
        } else {
            CK(cudaEventRecord(s->ev[1], s->stream));
        }
        CK(cudaEventRecord(s->ev[2], s->stream));
        CK(cudaMemcpyAsync(&cur, s->d_ctr, sizeof cur, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        vlog("filter passes done", (int)r);
        trace("filter passes done");
        {
            float t;
            if (brc < 0) {
                cudaEventElapsedTime(&t, s->ev[0], s->ev[1]); ms_fill += t;
                cudaEventElapsedTime(&t, s->ev[1], s->ev[2]); ms_query += t;
            } else {
                ms_bin += b_bin; ms_fill += b_fill; ms_query += b_query;
            }
        }
        if (merge_insert && r + 1 < s->rounds_eff) continue;   // exact pass after the last sub-round
        uint64_t marks_r = cur.marks - prev.marks;

        // exact set of this round's candidates, sized from the HyperLogLog estimate of their number
        std::vector<uint32_t> hll(1u << kHllBits);
        CK(cudaMemcpyAsync(hll.data(), s->d_hll, 4u << kHllBits, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        vlog("sketch read", (int)r);
        double est = hll_estimate(hll);
        if (est > (double)marks_r) est = (double)marks_r;
        uint32_t lg = std::max<uint32_t>(ceil_log2((uint64_t)(est * 1.15 * 2.0) + 64), 10);
        if (const char* e = getenv("TPC_TABLE_SHRINK")) lg = std::max<int>(4, (int)lg - atoi(e));  // tests: force the grow-and-redo path
        for (;;) {
            uint64_t need = sizeof(Slot) << lg;
            if (s->d_T && (s->T_in_scratch || need > s->T_bytes)) {
                if (!s->T_in_scratch) CK(dev_free(s->d_T, s->stream));
                s->d_T = nullptr; s->T_bytes = 0; s->T_in_scratch = false;
            }
            if (!s->d_T) {
                // the record waves of this round are spent (pipelined rounds: only this round's half is, unless
                // this is the last round)
                uint32_t* spent = s->d_bin_rec;
                uint64_t spent_bytes = s->bin_rec_bytes;
                if (s->pipe && r + 1 < s->rounds_eff) { spent = s->pipe_view[r & 1].rec; spent_bytes = s->bin_rec_bytes / 2; }
                if (s->d_bin_rec && need <= spent_bytes) {
                    s->d_T = reinterpret_cast<Slot*>(spent);
                    s->T_bytes = spent_bytes;
                    s->T_in_scratch = true;
                } else {
                    if (need > available_bytes(s->device)) return set_error("candidate table of 2^%u slots does not fit in device memory", lg);
                    CK(dev_alloc(&s->d_T, need, s->stream));
                    s->T_bytes = need;
                }
            }
            s->T_log2 = lg;
            CK(cudaEventRecord(s->ev[10], s->stream));
            CK(cudaMemsetAsync(s->d_T, 0, need, s->stream));
            // ownership planes (valid for every tile once a binned round has run) pick this round's marks
            OwnPlanes iop = s->own;
            iop.id = kp.part - s->prm.shard_index * s->rounds_eff + 1;
            const bool planes_ok = !merge_insert && brc == 0 && s->own_shared && kp.nparts > 1 && s->own_done_tiles >= s->ntiles;
            KParams kpi = kp;
            if (merge_insert) kpi.nparts = 1;   // every mark in the mask belongs to this pass
            // sparse marks that all went through the binned query kernels: insert from the mark list; else walk the mask
            const uint64_t list_cap = (uint64_t)s->marklist_regions * s->marklist_region_cap;
            const bool force_list = getenv("TPC_MARK_LIST_FORCE") != nullptr;   // (tests: the overflow -> redo path)
            const bool from_list = s->marklist_valid && ((marks_r * 10 <= list_cap * 8 && marks_r * 128 <= s->g.npos) || force_list);
            if (from_list) CK(W_DISPATCH(s, insert_list(lc, s->g, s->mark_list(true), kpi, TableView{s->d_T, lg, s->inline_keys()}, s->d_ctr)));
            else CK(W_DISPATCH(s, insert(lc, s->g, s->d_mask, kpi, 0, s->ntiles, TableView{s->d_T, lg, s->inline_keys()}, s->d_ctr,
                                         planes_ok ? &iop : nullptr)));
            CK(cudaEventRecord(s->ev[3], s->stream));
            CK(cudaMemcpyAsync(&cur, s->d_ctr, sizeof cur, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaStreamSynchronize(s->stream));
            vlog(from_list ? "insert (mark list) done" : "insert (mask) done", (int)r);
            const bool list_bad = cur.list_incomplete != prev.list_incomplete;   // a region had overflowed: redo from the mask
            if (list_bad) s->marklist_valid = false;
            if (!list_bad && cur.overflow == prev.overflow && (cur.distinct - prev.distinct) * 10 <= (7ull << lg)) break;
            // estimate too low (cannot happen within HLL's error bars, but stay exact): grow and redo
            Counters redo = cur;
            redo.distinct = prev.distinct; redo.overflow = prev.overflow; redo.list_incomplete = prev.list_incomplete;
            if (list_bad) --lg;   // same table size, other source
            CK(cudaMemcpyAsync(s->d_ctr, &redo, sizeof redo, cudaMemcpyHostToDevice, s->stream));
            cur = redo;
            ++lg;
        }
        TableView T{s->d_T, s->T_log2, s->inline_keys()};
        uint64_t distinct_r = cur.distinct - prev.distinct;
        if (s->local_count + distinct_r > s->local_cap) {
            // room for the remaining rounds too (they own equal shares), so the list is grown once
            uint64_t ncap = std::max<uint64_t>(s->local_count + distinct_r * (s->rounds_eff - r) * 9 / 8, 1024);
            unsigned long long* nl = nullptr;
            CK(dev_alloc(&nl, ncap * 8, s->stream));
            if (s->local_count) CK(cudaMemcpyAsync(nl, s->d_local, s->local_count * 8, cudaMemcpyDeviceToDevice, s->stream));
            CK(cudaStreamSynchronize(s->stream));
            if (s->d_local) CK(dev_free(s->d_local, s->stream));
            s->d_local = nl; s->local_cap = ncap;
        }
        CK(launch_classify(lc, T, s->prm.abundance, kp.count_occurrences, s->d_local, nullptr, s->local_cap, s->d_ctr));
        CK(cudaEventRecord(s->ev[4], s->stream));
        CK(cudaMemcpyAsync(&cur, s->d_ctr, sizeof cur, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        vlog("classify done", (int)r);
        s->local_count = cur.junctions;
        float t;
        cudaEventElapsedTime(&t, s->ev[10], s->ev[3]); ms_insert += t;
        cudaEventElapsedTime(&t, s->ev[3], s->ev[4]); ms_classify += t;
        prev = cur;
        // the table is only needed inside a round (h:337-338: per-round OccurenceSet); the next round's
        // record wave wants the memory
        if (s->T_in_scratch) { s->d_T = nullptr; s->T_bytes = 0; s->T_in_scratch = false; }
    }
    if (s->d_T) { CK(dev_free(s->d_T, s->stream)); s->d_T = nullptr; s->T_bytes = 0; }
    if (int rc = binned_release(s)) return rc;
    if (s->d_own_extra && !s->bin_keep) { CK(dev_free(s->d_own_extra, s->stream)); s->d_own_extra = nullptr; s->own_extra_bytes = 0; }
    s->st.candidate_marks = cur.marks;
    s->st.candidate_kmers = cur.distinct;
    s->st.filter_edges_set = cur.filter_new;
    s->st.ms_bin = ms_bin; s->st.ms_fill = ms_fill; s->st.ms_query = ms_query; s->st.ms_insert = ms_insert; s->st.ms_classify = ms_classify;
    s->have_candidates = true;
    s->have_index = false;
    vlog("scratch released", -1);
    trace("find_candidates: done");
    return 0;
}

int tpc_session_local_junctions(tpc_session* s, const uint64_t** dev_words, uint64_t* count) {
    if (!s || !s->have_candidates) return set_error("find_candidates has not run");
    if (dev_words) *dev_words = (const uint64_t*)s->d_local;
    if (count) *count = s->local_count;
    return 0;
}

int tpc_session_set_junctions(tpc_session* s, const uint64_t* dev_words_all, uint64_t n) {
    if (!s || !s->g.codes) return set_error("no genome set");
    if (int wrc = wait_genome(s, s->ntiles)) return wrc;
    LaunchCtx lc = s->lctx();
    WallTimer wall(&s->st.ms_wall_index);
    CK(cudaEventRecord(s->ev[5], s->stream));
    if (s->d_sorted) { CK(dev_free(s->d_sorted, s->stream)); s->d_sorted = nullptr; }
    if (s->d_J) { CK(dev_free(s->d_J, s->stream)); s->d_J = nullptr; }
    CK(dev_alloc(&s->d_sorted, std::max<uint64_t>(n, 1) * 8, s->stream));
    if (n) {
        // ids = rank of the first occurrence position: plain library radix sort of <= J 40-bit keys
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, (const unsigned long long*)dev_words_all, s->d_sorted, n, 0, kPosBits, s->stream));
        if (tmp > s->sort_tmp_bytes) {
            if (s->d_sort_tmp) CK(dev_free(s->d_sort_tmp, s->stream));
            CK(dev_alloc(&s->d_sort_tmp, tmp, s->stream));
            s->sort_tmp_bytes = tmp;
        }
        CK(cub::DeviceRadixSort::SortKeys(s->d_sort_tmp, tmp, (const unsigned long long*)dev_words_all, s->d_sorted, n, 0, kPosBits, s->stream));
    }
    s->J_log2 = std::max<uint32_t>(ceil_log2(n * 2 + 16), 6);
    CK(dev_alloc(&s->d_J, sizeof(Slot) << s->J_log2, s->stream));
    CK(cudaMemsetAsync(s->d_J, 0, sizeof(Slot) << s->J_log2, s->stream));
    CK(W_DISPATCH(s, build_index(lc, s->g, s->d_sorted, n, s->kparams(0), TableView{s->d_J, s->J_log2, s->inline_keys()})));
    CK(cudaEventRecord(s->ev[6], s->stream));
    s->J_count = n;
    s->st.junctions = n;
    s->have_index = true;
    s->have_count = false;
    return 0;
}

int tpc_session_candidate_mask(tpc_session* s, uint32_t** dev_mask, uint64_t* n_words) {
    if (s && s->windowed) return set_error("a windowed session keeps no candidate mask");
    if (!s || !s->d_mask) return set_error("find_candidates has not run");
    if (dev_mask) *dev_mask = s->d_mask;
    if (n_words) *n_words = padded_mask_words(s);
    return 0;
}

int tpc_session_emit_count(tpc_session* s, uint64_t pos_begin, uint64_t pos_end, uint64_t* n_records, uint64_t* n_stubs) {
    if (!s || !s->have_index || !s->d_mask) return set_error("set_junctions has not run");
    if (pos_end > s->g.npos) pos_end = s->g.npos;
    if (pos_begin > pos_end) pos_begin = pos_end;
    if (pos_begin % kTilePos && pos_begin != s->g.npos) return set_error("emit slice must start at a multiple of %d", kTilePos);
    if (pos_end % kTilePos && pos_end != s->g.npos) return set_error("emit slice must end at a multiple of %d", kTilePos);
    LaunchCtx lc = s->lctx();
    WallTimer wall(&s->st.ms_wall_emit);
    uint64_t tb = pos_begin / kTilePos, te = (pos_end + kTilePos - 1) / kTilePos;
    if (pos_begin >= pos_end) te = tb;   // empty slice (more shards than tiles)
    uint64_t nt = te - tb;
    if (nt + 1 > s->tile_cap) {
        for (void* p : {(void*)s->d_tile_rec, (void*)s->d_tile_stub, (void*)s->d_scan_scratch})
            if (p) CK(dev_free(p, s->stream));
        s->d_tile_rec = s->d_tile_stub = s->d_scan_scratch = nullptr;
        CK(dev_alloc(&s->d_tile_rec, (nt + 1) * 8, s->stream));
        CK(dev_alloc(&s->d_tile_stub, (nt + 1) * 8, s->stream));
        CK(dev_alloc(&s->d_scan_scratch, scan_scratch_items(nt + 1) * 8, s->stream));
        s->tile_cap = nt + 1;
    }
    CK(cudaEventRecord(s->ev[7], s->stream));
    if (te > tb) CK(cudaMemsetAsync(s->d_stubmask + tb * kTileThreads, 0, (te - tb) * kTileThreads * 4, s->stream));   // (only the slice is read)
    CK(cudaMemsetAsync(s->d_tile_rec + nt, 0, 8, s->stream));
    CK(cudaMemsetAsync(s->d_tile_stub + nt, 0, 8, s->stream));
    KParams kp = s->kparams(0);
    TableView J{s->d_J, s->J_log2, s->inline_keys()};
    CK(W_DISPATCH(s, ends(lc, s->g, s->rtable(), kp, J, s->d_stubmask, pos_begin, pos_end)));
    CK(W_DISPATCH(s, emit_count(lc, s->g, s->d_mask, s->d_stubmask, kp, J, tb, te, s->d_tile_rec, s->d_tile_stub)));
    CK(launch_scan_exclusive(lc, s->d_tile_rec, nt + 1, s->d_scan_scratch));
    CK(launch_scan_exclusive(lc, s->d_tile_stub, nt + 1, s->d_scan_scratch));
    unsigned long long tot[2];
    CK(cudaMemcpyAsync(&tot[0], s->d_tile_rec + nt, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(&tot[1], s->d_tile_stub + nt, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->slice_tile_begin = tb; s->slice_tile_end = te;
    s->slice_pos_begin = pos_begin; s->slice_pos_end = pos_end;
    s->slice_records = tot[0]; s->slice_stubs = tot[1];
    s->have_count = true;
    if (n_records) *n_records = tot[0];
    if (n_stubs) *n_stubs = tot[1];
    return 0;
}

// index of the last emitting record (len >= k) whose first position is < pos, or 0
static uint64_t emit_prev_at(const tpc_session* s, uint64_t pos) {
    uint64_t i = std::lower_bound(s->rec_start.begin(), s->rec_start.end(), pos) - s->rec_start.begin();
    return s->emit_prev[i];
}

int tpc_session_emit_write(tpc_session* s, uint64_t records_before, uint64_t stubs_before, uint8_t* dev_out,
                           uint64_t out_capacity, uint64_t* image_offset, uint64_t* image_bytes) {
    if (!s || !s->have_count) return set_error("emit_count has not run");
    LaunchCtx lc = s->lctx();
    WallTimer wall(&s->st.ms_wall_emit);
    uint64_t unit_base = records_before + emit_prev_at(s, s->slice_pos_begin);
    uint64_t units = s->slice_records + emit_prev_at(s, s->slice_pos_end) - emit_prev_at(s, s->slice_pos_begin);
    if (image_offset) *image_offset = unit_base * 12;
    if (image_bytes) *image_bytes = units * 12;
    if (units * 12 > out_capacity) {
        set_error("output buffer too small: need %llu bytes", (unsigned long long)(units * 12));
        return 2;
    }
    if (((uintptr_t)dev_out) & 3) return set_error("output buffer must be 4-byte aligned");
    KParams kp = s->kparams(0);
    TableView J{s->d_J, s->J_log2, s->inline_keys()};
    CK(W_DISPATCH(s, emit_write(lc, s->g, s->d_mask, s->d_stubmask, kp, J, s->rtable(), s->slice_tile_begin, s->slice_tile_end,
                                s->d_tile_rec, s->d_tile_stub, records_before, stubs_before, unit_base,
                                s->J_count + TPC_STUB_ID_OFFSET, (uint32_t*)dev_out, units)));
    CK(cudaEventRecord(s->ev[8], s->stream));
    s->st.occurrences = s->slice_records;
    s->st.stubs = s->slice_stubs;
    s->st.out_bytes = units * 12;
    return 0;
}

int tpc_session_get_id(tpc_session* s, const char* kmer, int64_t* id) {
    if (!s || !s->have_index || !kmer || !id) return set_error("bad argument");
    uint32_t k = s->prm.k;
    *id = TPC_INVALID_VERTEX;
    if (strlen(kmer) != k) return 0;
    uint64_t words[(TPC_MAX_K + 31) / 32] = {};
    for (uint32_t j = 0; j < k; ++j) {
        uint64_t c;
        switch (kmer[j]) {
            case 'A': case 'a': c = 0; break;
            case 'C': case 'c': c = 1; break;
            case 'G': case 'g': c = 2; break;
            case 'T': case 't': c = 3; break;
            default: return 0;  // not a definite k-mer: never a junction
        }
        words[j / 32] |= c << (2 * (j % 32));
    }
    LaunchCtx lc = s->lctx();
    CK(W_DISPATCH(s, get_id(lc, s->g, TableView{s->d_J, s->J_log2, s->inline_keys()}, s->kparams(0), words, s->d_id)));
    long long v = 0;
    CK(cudaMemcpyAsync(&v, s->d_id, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *id = v ? (int64_t)v : TPC_INVALID_VERTEX;
    return 0;
}

int tpc_session_stats(tpc_session* s, tpc_stats* out) {
    if (!s || !out) return set_error("null argument");
    CK(cudaStreamSynchronize(s->stream));
    float t = 0;
    if (s->have_index && cudaEventElapsedTime(&t, s->ev[5], s->ev[6]) == cudaSuccess) s->st.ms_index = t;
    if (s->st.out_bytes && cudaEventElapsedTime(&t, s->ev[7], s->ev[8]) == cudaSuccess) s->st.ms_emit = t;
    s->st.ms_total = s->st.ms_bin + s->st.ms_fill + s->st.ms_query + s->st.ms_insert + s->st.ms_classify + s->st.ms_index + s->st.ms_emit;
    s->st.kernel_launches = s->launches;
    s->st.skew_rebins = s->skew_rebins;
    s->st.h2d_bytes = s->h2d_bytes;
    *out = s->st;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// level 2: packed genome in host memory -> image in host memory (single GPU, unsharded)
// ---------------------------------------------------------------------------------------------
int tpc_session_run_to_count(tpc_session* s, uint64_t* image_bytes) {
    if (int rc = tpc_session_find_candidates(s)) return rc;
    const uint64_t* words = nullptr;
    uint64_t n = 0;
    if (int rc = tpc_session_local_junctions(s, &words, &n)) return rc;
    if (int rc = tpc_session_set_junctions(s, words, n)) return rc;
    uint64_t nrec = 0, nstub = 0;
    if (int rc = tpc_session_emit_count(s, 0, s->g.npos, &nrec, &nstub)) return rc;
    if (image_bytes) *image_bytes = (nrec + s->emit_prev[s->rec_start.size()]) * 12;
    return 0;
}

// Emit the counted slice in `parts` tile ranges; after each range `on_part(byte_lo, byte_hi)` is
// called with the image bytes that are now final on the device (the caller overlaps their transfer
// with the emission of the next range).
static int emit_write_parts(tpc_session* s, uint8_t* d_out, uint64_t image_bytes, uint32_t parts,
                            const std::function<int(uint64_t, uint64_t)>& on_part) {
    if (!s->have_count) return set_error("emit_count has not run");
    LaunchCtx lc = s->lctx();
    const uint64_t tb = s->slice_tile_begin, te = s->slice_tile_end, nt = te - tb;
    const uint64_t unit_base = emit_prev_at(s, s->slice_pos_begin);
    const uint64_t units = s->slice_records + emit_prev_at(s, s->slice_pos_end) - unit_base;
    if (units * 12 > image_bytes) return set_error("output buffer too small: need %llu bytes", (unsigned long long)(units * 12));
    parts = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(parts, nt));
    // record prefix at the part boundaries (tiny device->host reads of the scanned tile counts)
    std::vector<uint64_t> bt(parts + 1);
    std::vector<unsigned long long> pref(parts + 1, 0);
    for (uint32_t i = 0; i <= parts; ++i) {
        bt[i] = tb + nt * i / parts;
        CK(cudaMemcpyAsync(&pref[i], s->d_tile_rec + (bt[i] - tb), 8, cudaMemcpyDeviceToHost, s->stream));
    }
    CK(cudaStreamSynchronize(s->stream));
    KParams kp = s->kparams(0);
    TableView J{s->d_J, s->J_log2, s->inline_keys()};
    uint64_t byte_lo = 0;
    for (uint32_t i = 0; i < parts; ++i) {
        CK(W_DISPATCH(s, emit_write(lc, s->g, s->d_mask, s->d_stubmask, kp, J, s->rtable(), bt[i], bt[i + 1],
                                    s->d_tile_rec + (bt[i] - tb), s->d_tile_stub + (bt[i] - tb), 0, 0, unit_base,
                                    s->J_count + TPC_STUB_ID_OFFSET, (uint32_t*)d_out, units)));
        uint64_t pos_hi = std::min<uint64_t>(bt[i + 1] * kTilePos, s->slice_pos_end);
        uint64_t byte_hi = i + 1 == parts ? units * 12 : (pref[i + 1] + emit_prev_at(s, pos_hi) - unit_base) * 12;
        if (int rc = on_part(byte_lo, byte_hi)) return rc;
        byte_lo = byte_hi;
    }
    CK(cudaEventRecord(s->ev[8], s->stream));
    s->st.occurrences = s->slice_records;
    s->st.stubs = s->slice_stubs;
    s->st.out_bytes = units * 12;
    return 0;
}

int tpc_session_write_host(tpc_session* s, uint8_t* out_image, uint64_t image_bytes) {
    uint8_t* d_out = nullptr;
    CK(dev_alloc(&d_out, std::max<uint64_t>(image_bytes, 16), s->stream));
    DevBuf out_holder;
    out_holder.p = d_out; out_holder.st = s->stream;
    if (!s->copy_stream) CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    Events evs(1, cudaEventDisableTiming);
    cudaEvent_t done = evs[0];
    // device->host copy of part i (copy stream) overlaps the emission of part i+1 (compute stream)
    int rc = emit_write_parts(s, d_out, image_bytes, 8, [&](uint64_t lo, uint64_t hi) -> int {
        if (hi <= lo) return 0;
        CK(cudaEventRecord(done, s->stream));
        CK(cudaStreamWaitEvent(s->copy_stream, done, 0));
        CK(cudaMemcpyAsync(out_image + lo, d_out + lo, hi - lo, cudaMemcpyDeviceToHost, s->copy_stream));
        return 0;
    });
    cudaError_t e = cudaStreamSynchronize(s->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess && rc == 0) rc = set_error("CUDA error %s copying the image", cudaGetErrorName(e));
    return rc;
}

int tpc_session_write_stream(tpc_session* s, uint64_t image_bytes, tpc_chunk_sink sink, void* ctx) {
    uint8_t* d_out = nullptr;
    CK(dev_alloc(&d_out, std::max<uint64_t>(image_bytes, 16), s->stream));
    DevBuf out_holder;
    out_holder.p = d_out; out_holder.st = s->stream;
    uint64_t off = 0, bytes = 0;
    int rc = tpc_session_emit_write(s, 0, 0, d_out, image_bytes, &off, &bytes);
    const uint64_t kChunk = 64ull << 20;
    PinnedBuf pin[2];
    uint8_t* stage[2] = {nullptr, nullptr};
    Events evs(2, cudaEventDisableTiming);
    cudaEvent_t ev[2] = {evs[0], evs[1]};
    if (rc == 0 && bytes) {
        for (int i = 0; i < 2 && rc == 0; ++i) {
            if (cudaMallocHost(&pin[i].p, std::min(kChunk, bytes)) != cudaSuccess) rc = set_error("out of (pinned) host memory");
            stage[i] = (uint8_t*)pin[i].p;
        }
        uint64_t nchunks = (bytes + kChunk - 1) / kChunk;
        auto issue = [&](uint64_t c) {
            uint64_t lo = c * kChunk, n = std::min(kChunk, bytes - lo);
            cudaMemcpyAsync(stage[c & 1], d_out + lo, n, cudaMemcpyDeviceToHost, s->stream);
            cudaEventRecord(ev[c & 1], s->stream);
        };
        if (rc == 0) issue(0);
        for (uint64_t c = 0; c < nchunks && rc == 0; ++c) {
            if (cudaEventSynchronize(ev[c & 1]) != cudaSuccess) { rc = set_error("device to host copy failed"); break; }
            if (c + 1 < nchunks) issue(c + 1);   // overlaps the sink below
            rc = sink(ctx, stage[c & 1], std::min(kChunk, bytes - c * kChunk));
        }
    }
    cudaStreamSynchronize(s->stream);
    return rc;
}

static int host_windowed_run(tpc_session* s, const tpc_genome* g, uint64_t window_tiles, uint8_t* out_image, uint64_t out_capacity,
                             uint64_t* out_bytes);

int tpc_junctions_host(const tpc_params* params, const tpc_genome* host_genome, uint8_t* out_image, uint64_t out_capacity,
                       uint64_t* out_bytes, tpc_stats* stats) {
    if (!params || !host_genome) return set_error("null argument");
    tpc_params p = *params;
    p.shard_index = 0;
    p.shard_count = 1;
    tpc_session* s = nullptr;
    if (int rc = tpc_session_create(&p, nullptr, &s)) return rc;
    // Inputs that do not fit HBM beside the filter (genome 0.375 B + candidate / stub masks 0.25 B per position), or
    // TPC_WINDOW_TILES=<tiles per window>: the position-windowed driver (tpc_windowed.inl)
    uint64_t window_tiles = 0;
    if (const char* e = getenv("TPC_WINDOW_TILES")) window_tiles = (uint64_t)std::max(0ll, atoll(e));
    else {
        const double resident = 0.75 * (double)host_genome->n_positions + (double)((1ull << std::max<uint32_t>(p.filter_bits, 9u)) / 8);
        if (resident > 0.85 * (double)available_bytes(s->device)) window_tiles = 1u << 17;   // 2^30 positions per window
    }
    if (window_tiles) {
        int rc = host_windowed_run(s, host_genome, window_tiles, out_image, out_capacity, out_bytes);
        if (stats) tpc_session_stats(s, stats);
        tpc_session_destroy(s);
        return rc;
    }
    const bool verbose = getenv("TPC_VERBOSE") != nullptr;
    const double t_start = now_ms();
    auto vlog = [&](const char* what) {
        if (verbose) fprintf(stderr, "[tpc junctions_host] +%9.3f ms  %s\n", now_ms() - t_start, what);
    };
    vlog("session created");
    int rc = tpc_session_set_genome_host(s, host_genome);
    vlog("genome upload enqueued");
    uint64_t bytes = 0;
    if (rc == 0) rc = tpc_session_run_to_count(s, &bytes);
    vlog("candidates, index, emit count done");
    if (out_bytes) *out_bytes = bytes;
    if (rc == 0 && bytes > out_capacity) {
        set_error("output buffer too small: need %llu bytes", (unsigned long long)bytes);
        rc = 2;
    }
    if (rc == 0) rc = tpc_session_write_host(s, out_image, bytes);
    vlog("image written to the host buffer");
    if (stats) tpc_session_stats(s, stats);
    tpc_session_destroy(s);
    vlog("session destroyed");
    return rc;
}

}  // extern "C"

#include "tpc_windowed.inl"

// level 2, one GPU, windowed: packed genome in host memory -> image in host memory
namespace {
struct HostImageSink {
    uint8_t* image;
    uint64_t capacity, written;
};
int host_image_sink(void* ctx, const uint8_t* dev_bytes, uint64_t image_offset, uint64_t nbytes, cudaStream_t stream) {
    HostImageSink* k = static_cast<HostImageSink*>(ctx);
    k->written = std::max(k->written, image_offset + nbytes);
    if (image_offset + nbytes > k->capacity || nbytes == 0) return 0;   // (too small: keep counting, report the size at the end)
    return cudaMemcpyAsync(k->image + image_offset, dev_bytes, nbytes, cudaMemcpyDeviceToHost, stream) == cudaSuccess
               ? 0 : tpc::set_error("device to host copy of the image failed");
}
}  // namespace

static int host_windowed_run(tpc_session* s, const tpc_genome* g, uint64_t window_tiles, uint8_t* out_image, uint64_t out_capacity,
                             uint64_t* out_bytes) {
    HostWindowProvider prov(g->codes, g->n_mask, g->n_positions, window_tiles, 0, 1, nullptr);
    if (int rc = prov.init()) return rc;
    if (int rc = tpc_session_set_genome_windowed(s, g->n_positions, g->rec_start, g->rec_len, g->n_records, window_tiles, &prov)) return rc;
    if (int rc = tpc_session_find_candidates(s)) return rc;
    const uint64_t *words = nullptr, *keys = nullptr;
    uint64_t n = 0;
    if (int rc = tpc_session_local_junctions(s, &words, &n)) return rc;
    if (int rc = tpc_session_local_junction_keys(s, &keys)) return rc;
    if (int rc = tpc_session_set_junctions_keyed(s, words, keys, n)) return rc;
    HostImageSink sink{out_image, out_image ? out_capacity : 0, 0};
    uint64_t nrec = 0, nstub = 0;
    if (int rc = tpc_session_emit_windowed(s, 0, g->n_positions, 1, 0, 0, host_image_sink, &sink, &nrec, &nstub)) return rc;
    const uint64_t bytes = (nrec + s->emit_prev[s->rec_start.size()]) * 12;
    if (out_bytes) *out_bytes = bytes;
    s->st.out_bytes = bytes;
    s->wp = nullptr;   // (the provider dies with this frame)
    if (bytes > out_capacity) {
        tpc::set_error("output buffer too small: need %llu bytes", (unsigned long long)bytes);
        return 2;
    }
    return 0;
}
