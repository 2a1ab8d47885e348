// tpc_multi.cpp -- the path on N GPUs of ONE process: one host thread per GPU, one session (hash-range shard)
// per GPU, NCCL over NVLink between the stages (DESIGN.md "Multi-GPU"):
//
//   0. genome   host buffers : every GPU uploads 1/N of each chunk over its own PCIe link, ncclAllGather in place
//               GPU 0 holds it: chunked ncclBroadcast (tpc_build: FASTA was parsed once and packed on GPU 0)
//               -- either way the sessions start on the chunks that have arrived (tpc_session_add_genome_event)
//   1. ncclBroadcast x N (an all-gather of variable-length lists) of the shards' junction words -> identical ids
//   2. ncclReduceScatter (sum == OR: the shards' masks are disjoint) of the candidate masks into the position slices
//   3. prefix of (records, stubs) over the slices (host memory: the threads share an address space), ordered emit,
//      every GPU writes its slice of the image (host buffer, or pwrite() into the output file)
//
// This replaces what the reference does with `threads` workers over one shared filter (vertexenumerator.h:122-466);
// it is what tpc_build / the twopaco CLI run when more than one GPU is visible.  NCCL is loaded with dlopen so that
// libtwopaco_b200.so has no link-time dependency on it (single-GPU use never touches it).
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>

#include <cuda_runtime_api.h>
#include <nccl.h>

#include "tpc_internal.h"
#include "tpc_multi.h"
#include "tpc_window_provider.h"

using tpc::set_error;

namespace tpc {

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen
// ---------------------------------------------------------------------------------------------
struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    static Nccl& get() {
        static Nccl n;
        static std::once_flag once;
        std::call_once(once, [] {
            const char* names[] = {getenv("TPC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
            for (const char* nm : names) {
                if (!nm || !*nm) continue;
                n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
                if (n.handle) break;
            }
            if (!n.handle) {
                n.error = std::string("NCCL is not available (") + (dlerror() ? dlerror() : "libnccl.so.2 not found") + ")";
                return;
            }
            bool ok = true;
            auto sym = [&](const char* name) -> void* {
                void* p = dlsym(n.handle, name);
                if (!p) { ok = false; n.error = std::string("NCCL symbol missing: ") + name; }
                return p;
            };
            n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
            n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
            n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
            n.ReduceScatter = (decltype(n.ReduceScatter))sym("ncclReduceScatter");
            n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
            n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
            n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
            n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
            if (!ok) { dlclose(n.handle); n.handle = nullptr; }
        });
        return n;
    }
};

namespace {

// reusable barrier for the worker threads (C++17 has none)
class HostBarrier {
public:
    explicit HostBarrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const int gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen != gen_; });
        }
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

constexpr uint64_t kTilePositions = 8192, kCodeWordsPerTile = 256, kMaskWordsPerTile = 128;

// Geometry of the chunked multi-GPU upload (same as twopaco_b200/dist.py ChunkPlan): the tiles are cut into at most
// n_chunks chunks of a multiple of `world` tiles; every chunk of both arrays is `world` equal parts, part r uploaded
// by GPU r; one in-place all-gather per chunk and array lands the chunk contiguously in every GPU's copy.  The last
// chunk also carries the arrays' read-ahead padding and is padded to a multiple of `world` words.
struct ChunkPlan {
    uint64_t tiles = 0;
    std::vector<uint64_t> tile_begin;
    struct Part { uint64_t start, part; };
    std::vector<Part> arr[2];   // [0] codes, [1] n_mask
    uint64_t total[2] = {0, 0};
    ChunkPlan(uint64_t n_positions, uint64_t code_words, uint64_t mask_words, uint32_t world, uint32_t n_chunks) {
        tiles = (n_positions + kTilePositions - 1) / kTilePositions;
        uint64_t per = std::max<uint64_t>(1, (tiles + n_chunks - 1) / std::max<uint32_t>(n_chunks, 1));
        per = (per + world - 1) / world * world;
        for (uint64_t t = 0; t < std::max<uint64_t>(tiles, 1); t += per) tile_begin.push_back(t);
        total[0] = code_words; total[1] = mask_words;
        const uint64_t wpt[2] = {kCodeWordsPerTile, kMaskWordsPerTile};
        for (int a = 0; a < 2; ++a)
            for (size_t c = 0; c < tile_begin.size(); ++c) {
                const uint64_t start = tile_begin[c] * wpt[a];
                const uint64_t end = c + 1 == tile_begin.size() ? total[a] : tile_begin[c + 1] * wpt[a];
                arr[a].push_back(Part{start, (end - start + world - 1) / world});
            }
    }
    uint64_t device_words(int a, uint32_t world) const { return arr[a].back().start + arr[a].back().part * world; }
};

struct Shared {
    int n = 0;
    std::vector<int> dev;
    std::vector<ncclComm_t> comm;
    std::vector<int> rc;
    std::vector<std::string> err;
    std::vector<uint64_t> jcount, nrec, nstub, slice_off, slice_bytes;
    std::vector<std::array<uint64_t, 2>> digest;
    std::vector<tpc_stats> stats;
    std::vector<tpc_session*> session;
    std::vector<char> flag;
    HostBarrier bar;
    // logical OR of the shards' flags (every shard's thread calls it)
    bool any(int r, bool mine) {
        flag[r] = mine ? 1 : 0;
        bar.wait();
        bool a = false;
        for (int i = 0; i < n; ++i) a = a || flag[i];
        bar.wait();
        return a;
    }
    explicit Shared(int n_) : n(n_), dev(n_), comm(n_, nullptr), rc(n_, 0), err(n_), jcount(n_, 0), nrec(n_, 0), nstub(n_, 0),
                              slice_off(n_, 0), slice_bytes(n_, 0), digest(n_, std::array<uint64_t, 2>{0, 0}), stats(n_), session(n_, nullptr), flag(n_, 0), bar(n_) {}
    // all threads call this with their own status; returns true when every thread is fine
    bool sync_ok(int r, int my_rc) {
        if (my_rc != 0 && rc[r] == 0) { rc[r] = my_rc; err[r] = tpc::last_error(); }
        bar.wait();
        bool ok = true;
        for (int i = 0; i < n; ++i) ok = ok && rc[i] == 0;
        bar.wait();   // nobody changes rc[] before everybody has read it
        return ok;
    }
};

#define CKM(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)
#define CKN(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t e_ = (call);                                                                          \
        if (e_ != ncclSuccess)                                                                             \
            return set_error("NCCL error at %s:%d (%s)", __FILE__, __LINE__, Nccl::get().GetErrorString(e_)); \
    } while (0)

uint64_t slice_cut(uint64_t n_positions, uint32_t world, uint32_t r) {   // == dist.position_cuts
    const uint64_t tiles = (n_positions + kTilePositions - 1) / kTilePositions, chunk = (tiles + world - 1) / world;
    return r >= world ? n_positions : std::min<uint64_t>(n_positions, (uint64_t)r * chunk * kTilePositions);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// the multi-GPU context: devices + communicators, created once, reused by every run
// ---------------------------------------------------------------------------------------------
struct MultiImpl {
    std::vector<int> dev;
    std::vector<ncclComm_t> comm;
    ~MultiImpl() {
        Nccl& nc = Nccl::get();
        for (size_t i = 0; i < comm.size(); ++i)
            if (comm[i] && nc.CommDestroy) {
                cudaSetDevice(dev[i]);
                nc.CommDestroy(comm[i]);
            }
    }
};

// where a rank's slice of the image goes
struct ImageSink {
    uint8_t* host_image = nullptr;   // level 2: bytes [offset, offset + n) of this buffer
    uint64_t host_capacity = 0;
    int fd = -1;                     // level 1: pwrite() at the offset
    bool digest_only = false;        // neither: the image never leaves the GPUs, only its digest does (Shared::digest)
};

struct GenomeSource {
    const tpc_genome* host = nullptr;          // packed genome in host memory (chunked 1/N upload + all-gather)
    const uint64_t* dev0_codes = nullptr;      // or: the packed genome already on GPU 0 (chunked broadcast)
    const uint64_t* dev0_nmask = nullptr;
    uint64_t n_positions = 0;
    const uint64_t *rec_start = nullptr, *rec_len = nullptr;
    uint64_t n_records = 0;
    uint64_t window_tiles = 0;                 // > 0: position-windowed run (host source only; tpc_windowed.inl)
};

// a window's part of the image (windowed runs): to the host image, the file, or into the digest
struct WindowSinkCtx {
    const ImageSink* sink;
    uint64_t digest[2] = {0, 0};
    uint64_t end = 0;
    void* pin = nullptr;
    uint64_t pin_cap = 0;
};
static int window_sink(void* ctx, const uint8_t* dev_bytes, uint64_t image_offset, uint64_t nbytes, cudaStream_t stream) {
    WindowSinkCtx* w = static_cast<WindowSinkCtx*>(ctx);
    w->end = std::max(w->end, image_offset + nbytes);
    if (nbytes == 0) return 0;
    if (w->sink->digest_only) {
        uint64_t d[2];
        if (int rc = tpc_image_digest_device(dev_bytes, nbytes, image_offset, stream, d)) return rc;
        w->digest[0] += d[0]; w->digest[1] += d[1];
        return 0;
    }
    if (w->sink->host_image) {
        if (image_offset + nbytes > w->sink->host_capacity) return 0;   // (too small: the size is reported at the end)
        return cudaMemcpyAsync(w->sink->host_image + image_offset, dev_bytes, nbytes, cudaMemcpyDeviceToHost, stream) == cudaSuccess
                   ? 0 : set_error("device to host copy of the image failed");
    }
    if (w->sink->fd >= 0) {
        if (w->pin_cap < nbytes) {
            if (w->pin) cudaFreeHost(w->pin);
            w->pin = nullptr;
            if (cudaMallocHost(&w->pin, nbytes + nbytes / 4) != cudaSuccess) return set_error("out of pinned host memory");
            w->pin_cap = nbytes + nbytes / 4;
        }
        if (cudaMemcpyAsync(w->pin, dev_bytes, nbytes, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess)
            return set_error("device to host copy of the image failed");
        uint64_t done = 0;
        while (done < nbytes) {
            ssize_t n = pwrite(w->sink->fd, (const uint8_t*)w->pin + done, nbytes - done, (off_t)(image_offset + done));
            if (n <= 0) return set_error("Can't write to the output file");
            done += (uint64_t)n;
        }
    }
    return 0;
}

static int shard_worker_windowed(MultiImpl* mi, Shared* sh, int r, const tpc_params& base, const GenomeSource& src, const ImageSink& sink,
                                 uint64_t* image_bytes_out);

static int shard_worker(MultiImpl* mi, Shared* sh, int r, const tpc_params& base, const GenomeSource& src, const ImageSink& sink,
                        bool keep_session0, uint64_t* image_bytes_out) {
    if (src.window_tiles) return shard_worker_windowed(mi, sh, r, base, src, sink, image_bytes_out);
    Nccl& nc = Nccl::get();
    const int N = sh->n;
    int rc = 0;
    const bool verbose = getenv("TPC_VERBOSE") != nullptr && r == 0;
    const auto t_start = std::chrono::steady_clock::now();
    auto vlog = [&](const char* what) {
        if (verbose)
            fprintf(stderr, "[tpc multi, GPU 0 of %d] +%9.3f ms  %s\n", N,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(), what);
    };
    cudaStream_t st = nullptr, h2d = nullptr, ag = nullptr;
    uint64_t *d_codes = nullptr, *d_nmask = nullptr;
    unsigned long long* d_allj = nullptr;
    uint8_t* d_out = nullptr;
    void* pin[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_chunk, ev_tmp;
    tpc_session* s = nullptr;
    const bool own_genome = !(r == 0 && src.dev0_codes);
    const uint64_t cw = tpc_code_words(src.n_positions), mw = tpc_mask_words(src.n_positions);
    ChunkPlan plan(src.n_positions, cw, mw, (uint32_t)N, 16);

    auto body = [&]() -> int {
        CKM(cudaSetDevice(sh->dev[r]));
        CKM(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CKM(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
        CKM(cudaStreamCreateWithFlags(&ag, cudaStreamNonBlocking));
        // ---- 0. the whole packed genome on this GPU, chunk by chunk
        const uint64_t* codes = src.dev0_codes;
        const uint64_t* nmask = src.dev0_nmask;
        if (own_genome) {   // (stream-ordered: the device's pool keeps the memory between runs)
            CKM(cudaMallocAsync((void**)&d_codes, plan.device_words(0, N) * 8, st));
            CKM(cudaMallocAsync((void**)&d_nmask, plan.device_words(1, N) * 8, st));
            CKM(cudaStreamSynchronize(st));
            codes = d_codes; nmask = d_nmask;
        }
        vlog("streams + genome arrays ready");
        uint64_t* full[2] = {const_cast<uint64_t*>(codes), const_cast<uint64_t*>(nmask)};
        const uint64_t* host_arr[2] = {src.host ? src.host->codes : nullptr, src.host ? src.host->n_mask : nullptr};
        for (size_t c = 0; c < plan.tile_begin.size(); ++c) {
            for (int a = 0; a < 2; ++a) {
                const uint64_t start = plan.arr[a][c].start, part = plan.arr[a][c].part;
                if (src.host) {
                    // my part of the chunk over my PCIe link, then the in-place all-gather over NVLink
                    const uint64_t lo = std::min(plan.total[a], start + (uint64_t)r * part), hi = std::min(plan.total[a], lo + part);
                    if (hi > lo) CKM(cudaMemcpyAsync(full[a] + lo, host_arr[a] + lo, (hi - lo) * 8, cudaMemcpyHostToDevice, h2d));
                    if (hi < start + (uint64_t)(r + 1) * part)   // padding of the last chunk
                        CKM(cudaMemsetAsync(full[a] + hi, 0, (start + (uint64_t)(r + 1) * part - hi) * 8, h2d));
                    cudaEvent_t e;
                    CKM(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    ev_tmp.push_back(e);
                    CKM(cudaEventRecord(e, h2d));
                    CKM(cudaStreamWaitEvent(ag, e, 0));
                    CKN(nc.AllGather(full[a] + start + (uint64_t)r * part, full[a] + start, part, ncclUint64, sh->comm[r], ag));
                } else {
                    const uint64_t count = std::min(plan.total[a], start + part * N) - start;
                    CKN(nc.Broadcast(full[a] + start, full[a] + start, count, ncclUint64, 0, sh->comm[r], ag));
                }
            }
            cudaEvent_t e;
            CKM(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev_chunk.push_back(e);
            CKM(cudaEventRecord(e, ag));
        }
        // ---- the shard's session
        tpc_params prm = base;
        prm.shard_index = (uint32_t)r; prm.shard_count = (uint32_t)N;
        if (int e = tpc_session_create(&prm, st, &s)) return e;
        sh->session[r] = s;
        tpc_genome g{};
        g.codes = codes; g.n_mask = nmask; g.n_positions = src.n_positions;
        g.rec_start = src.rec_start; g.rec_len = src.rec_len; g.n_records = src.n_records;
        if (int e = tpc_session_set_genome_device(s, &g)) return e;
        if (own_genome)
            for (size_t c = 0; c < plan.tile_begin.size(); ++c)
                if (int e = tpc_session_add_genome_event(s, plan.tile_begin[c], ev_chunk[c])) return e;
        vlog("upload / all-gather enqueued, session ready");
        return tpc_session_find_candidates(s);
    };
    rc = body();
    vlog("find_candidates done");
    if (!sh->sync_ok(r, rc)) return rc;
    vlog("all shards done");

    // ---- 1. all-gather of the shards' junction words
    const uint64_t* words = nullptr;
    uint64_t nj = 0;
    rc = tpc_session_local_junctions(s, &words, &nj);
    sh->jcount[r] = nj;
    if (!sh->sync_ok(r, rc)) return rc;
    uint64_t total_j = 0, my_off = 0;
    for (int i = 0; i < N; ++i) { if (i < r) my_off += sh->jcount[i]; total_j += sh->jcount[i]; }
    auto exchange = [&]() -> int {
        CKM(cudaMallocAsync((void**)&d_allj, std::max<uint64_t>(total_j, 1) * 8, st));
        if (nj) CKM(cudaMemcpyAsync(d_allj + my_off, words, nj * 8, cudaMemcpyDeviceToDevice, st));
        CKN(nc.GroupStart());
        uint64_t off = 0;
        for (int i = 0; i < N; ++i) {
            if (sh->jcount[i]) CKN(nc.Broadcast(d_allj + off, d_allj + off, sh->jcount[i], ncclUint64, i, sh->comm[r], st));
            off += sh->jcount[i];
        }
        CKN(nc.GroupEnd());
        if (int e = tpc_session_set_junctions(s, (const uint64_t*)d_allj, total_j)) return e;
        // ---- 2. OR-reduce-scatter of the disjoint candidate masks straight into the position slices
        uint32_t* mask = nullptr;
        uint64_t mwords = 0;
        if (int e = tpc_session_candidate_mask(s, &mask, &mwords)) return e;
        const uint64_t chunk = mwords / N;
        CKN(nc.ReduceScatter(mask, mask + (uint64_t)r * chunk, chunk, ncclUint32, ncclSum, sh->comm[r], st));
        // ---- 3. position-sharded emit
        uint64_t nr = 0, ns = 0;
        if (int e = tpc_session_emit_count(s, slice_cut(src.n_positions, N, r), slice_cut(src.n_positions, N, r + 1), &nr, &ns)) return e;
        sh->nrec[r] = nr; sh->nstub[r] = ns;
        return 0;
    };
    rc = exchange();
    vlog("junction all-gather, index, mask reduce-scatter, emit count done");
    if (!sh->sync_ok(r, rc)) return rc;
    uint64_t rb = 0, sb = 0;
    for (int i = 0; i < r; ++i) { rb += sh->nrec[i]; sb += sh->nstub[i]; }
    auto emit = [&]() -> int {
        // records of the slice + the separators that may precede them (at most one per record of the input)
        const uint64_t cap = 12 * (sh->nrec[r] + src.n_records) + 16;
        CKM(cudaMallocAsync((void**)&d_out, cap, st));
        uint64_t off = 0, nb = 0;
        if (int e = tpc_session_emit_write(s, rb, sb, d_out, cap, &off, &nb)) return e;
        sh->slice_off[r] = off; sh->slice_bytes[r] = nb;
        if (sink.digest_only) {
            uint64_t d[2] = {0, 0};
            if (int e = tpc_image_digest_device(d_out, nb, off, st, d)) return e;
            sh->digest[r][0] = d[0]; sh->digest[r][1] = d[1];
        } else if (sink.host_image) {
            if (off + nb > sink.host_capacity) {
                set_error("output buffer too small: need at least %llu bytes", (unsigned long long)(off + nb));
                return 2;
            }
            if (nb) CKM(cudaMemcpyAsync(sink.host_image + off, d_out, nb, cudaMemcpyDeviceToHost, st));
            CKM(cudaStreamSynchronize(st));
        } else if (sink.fd >= 0) {
            // two pinned staging buffers: the device->host copy of piece i+1 overlaps the pwrite of piece i
            const uint64_t kPiece = 32ull << 20;
            cudaEvent_t e2[2];
            for (int i = 0; i < 2; ++i) {
                CKM(cudaMallocHost(&pin[i], std::max<uint64_t>(std::min(kPiece, nb), 16)));
                CKM(cudaEventCreateWithFlags(&e2[i], cudaEventDisableTiming));
                ev_tmp.push_back(e2[i]);
            }
            const uint64_t pieces = (nb + kPiece - 1) / kPiece;
            auto issue = [&](uint64_t p) -> int {
                const uint64_t lo = p * kPiece, n = std::min(kPiece, nb - lo);
                CKM(cudaMemcpyAsync(pin[p & 1], d_out + lo, n, cudaMemcpyDeviceToHost, st));
                CKM(cudaEventRecord(e2[p & 1], st));
                return 0;
            };
            if (pieces)
                if (int e = issue(0)) return e;
            for (uint64_t p = 0; p < pieces; ++p) {
                CKM(cudaEventSynchronize(e2[p & 1]));
                const uint64_t lo = p * kPiece, n = std::min(kPiece, nb - lo);
                const uint8_t* from = (const uint8_t*)pin[p & 1];
                uint64_t done = 0;
                while (done < n) {   // (written before the next copy into this buffer is issued)
                    ssize_t w = pwrite(sink.fd, from + done, n - done, (off_t)(off + lo + done));
                    if (w <= 0) return set_error("Can't write to the output file");
                    done += (uint64_t)w;
                }
                if (p + 1 < pieces)
                    if (int e = issue(p + 1)) return e;
            }
            CKM(cudaStreamSynchronize(st));
        } else {
            CKM(cudaStreamSynchronize(st));
        }
        return tpc_session_stats(s, &sh->stats[r]);
    };
    rc = emit();
    vlog("emit + copy-out done");
    if (r == N - 1 && image_bytes_out && rc == 0) *image_bytes_out = sh->slice_off[r] + sh->slice_bytes[r];
    const bool ok = sh->sync_ok(r, rc);

    // ---- release
    cudaSetDevice(sh->dev[r]);
    if (st) cudaStreamSynchronize(st);
    if (ag) cudaStreamSynchronize(ag);
    if (d_allj) cudaFreeAsync(d_allj, st);
    if (d_out) cudaFreeAsync(d_out, st);
    const bool keep = ok && keep_session0 && r == 0;
    if (s && !keep) { tpc_session_destroy(s); sh->session[r] = nullptr; }
    if (!keep) {
        if (d_codes) cudaFreeAsync(d_codes, st);
        if (d_nmask) cudaFreeAsync(d_nmask, st);
        if (st) cudaStreamSynchronize(st);
    }
    for (auto e : ev_chunk) cudaEventDestroy(e);
    for (auto e : ev_tmp) cudaEventDestroy(e);
    for (void* p : pin)
        if (p) cudaFreeHost(p);
    if (st && !keep) cudaStreamDestroy(st);   // (a kept session keeps using its stream)
    if (h2d) cudaStreamDestroy(h2d);
    if (ag) cudaStreamDestroy(ag);
    vlog("released");
    return rc;
}

static int run_shards(MultiImpl* mi, const tpc_params& prm, const GenomeSource& src, const ImageSink& sink, bool keep_session0,
                      uint64_t* image_bytes, tpc_stats* stats, tpc_session** session0, uint64_t* digest = nullptr) {
    const int N = (int)mi->dev.size();
    Shared sh(N);
    sh.dev = mi->dev;
    sh.comm = mi->comm;
    std::vector<std::thread> th;
    std::vector<int> rcs(N, 0);
    for (int r = 1; r < N; ++r)
        th.emplace_back([&, r] { rcs[r] = shard_worker(mi, &sh, r, prm, src, sink, keep_session0, image_bytes); });
    int prev = 0;
    cudaGetDevice(&prev);
    rcs[0] = shard_worker(mi, &sh, 0, prm, src, sink, keep_session0, image_bytes);
    for (auto& t : th) t.join();
    cudaSetDevice(prev);
    for (int r = 0; r < N; ++r)
        if (sh.rc[r]) {
            set_error("GPU %d: %s", mi->dev[r], sh.err[r].c_str());
            if (sh.session[0]) { tpc_session_destroy(sh.session[0]); sh.session[0] = nullptr; }
            return sh.rc[r];
        }
    if (stats) {   // the log counters of the whole run: sums over the shards; stage times: the slowest shard
        tpc_stats t = sh.stats[0];
        for (int r = 1; r < N; ++r) {
            const tpc_stats& o = sh.stats[r];
            t.candidate_marks += o.candidate_marks; t.candidate_kmers += o.candidate_kmers; t.filter_edges_set += o.filter_edges_set;
            t.occurrences += o.occurrences; t.stubs += o.stubs; t.out_bytes += o.out_bytes; t.kernel_launches += o.kernel_launches;
            float* a = &t.ms_bin;
            const float* b = &o.ms_bin;
            for (int j = 0; j < 8; ++j) a[j] = std::max(a[j], b[j]);   // ms_bin .. ms_total
            t.ms_bin_overlapped = std::max(t.ms_bin_overlapped, o.ms_bin_overlapped);
            t.ms_wall_candidates = std::max(t.ms_wall_candidates, o.ms_wall_candidates);
            t.ms_wall_index = std::max(t.ms_wall_index, o.ms_wall_index);
            t.ms_wall_emit = std::max(t.ms_wall_emit, o.ms_wall_emit);
        }
        *stats = t;
    }
    if (session0) *session0 = sh.session[0];
    if (digest) {
        digest[0] = digest[1] = 0;
        for (int r = 0; r < N; ++r) { digest[0] += sh.digest[r][0]; digest[1] += sh.digest[r][1]; }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// windowed shards (tpc_windowed.inl): the packed genome stays in host memory and streams through every GPU once per
// pass -- shared passes: 1/N of each window per GPU over its own PCIe link + ncclAllGather; the position-sharded emit:
// each GPU uploads the windows of its slice itself.  No candidate mask exists, so there is no mask exchange; the
// junction exchange carries (first position, key) pairs.
// ---------------------------------------------------------------------------------------------
static int shard_worker_windowed(MultiImpl* mi, Shared* sh, int r, const tpc_params& base, const GenomeSource& src, const ImageSink& sink,
                                 uint64_t* image_bytes_out) {
    Nccl& nc = Nccl::get();
    const int N = sh->n;
    const bool verbose = getenv("TPC_VERBOSE") != nullptr && r == 0;
    const auto t_start = std::chrono::steady_clock::now();
    auto vlog = [&](const char* what) {
        if (verbose)
            fprintf(stderr, "[tpc multi windowed, GPU 0 of %d] +%9.3f ms  %s\n", N,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(), what);
    };
    int rc = 0;
    cudaStream_t st = nullptr;
    tpc_session* s = nullptr;
    unsigned long long *d_allj = nullptr, *d_allk = nullptr;
    HostWindowProvider* prov = nullptr;
    WindowSinkCtx wctx;
    wctx.sink = &sink;
    auto body = [&]() -> int {
        CKM(cudaSetDevice(sh->dev[r]));
        CKM(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ncclComm_t comm = sh->comm[r];
        prov = new HostWindowProvider(src.host->codes, src.host->n_mask, src.n_positions, src.window_tiles, r, N,
                                      [comm, r, &nc](uint64_t* buffer, uint64_t part_words, cudaStream_t stream) -> int {
                                          ncclResult_t e = nc.AllGather(buffer + (uint64_t)r * part_words, buffer, part_words, ncclUint64, comm, stream);
                                          return e == ncclSuccess ? 0 : set_error("ncclAllGather of a genome window failed: %s", nc.GetErrorString(e));
                                      });
        if (int e = prov->init()) return e;
        prov->set_agree([sh, r](bool mine) { return sh->any(r, mine); });
        tpc_params prm = base;
        prm.shard_index = (uint32_t)r; prm.shard_count = (uint32_t)N;
        if (int e = tpc_session_create(&prm, st, &s)) return e;
        sh->session[r] = s;
        if (int e = tpc_session_set_genome_windowed(s, src.n_positions, src.rec_start, src.rec_len, src.n_records, src.window_tiles, prov)) return e;
        return tpc_session_find_candidates(s);
    };
    rc = body();
    vlog("find_candidates done");
    if (!sh->sync_ok(r, rc)) goto done;
    {
        const uint64_t *words = nullptr, *keys = nullptr;
        uint64_t nj = 0;
        rc = tpc_session_local_junctions(s, &words, &nj);
        if (rc == 0) rc = tpc_session_local_junction_keys(s, &keys);
        sh->jcount[r] = nj;
        if (!sh->sync_ok(r, rc)) goto done;
        uint64_t total_j = 0, my_off = 0;
        for (int i = 0; i < N; ++i) { if (i < r) my_off += sh->jcount[i]; total_j += sh->jcount[i]; }
        auto exchange = [&]() -> int {
            CKM(cudaMallocAsync((void**)&d_allj, std::max<uint64_t>(total_j, 1) * 8, st));
            CKM(cudaMallocAsync((void**)&d_allk, std::max<uint64_t>(total_j, 1) * 8, st));
            if (nj) {
                CKM(cudaMemcpyAsync(d_allj + my_off, words, nj * 8, cudaMemcpyDeviceToDevice, st));
                CKM(cudaMemcpyAsync(d_allk + my_off, keys, nj * 8, cudaMemcpyDeviceToDevice, st));
            }
            CKN(nc.GroupStart());
            uint64_t off = 0;
            for (int i = 0; i < N; ++i) {
                if (sh->jcount[i]) {
                    CKN(nc.Broadcast(d_allj + off, d_allj + off, sh->jcount[i], ncclUint64, i, sh->comm[r], st));
                    CKN(nc.Broadcast(d_allk + off, d_allk + off, sh->jcount[i], ncclUint64, i, sh->comm[r], st));
                }
                off += sh->jcount[i];
            }
            CKN(nc.GroupEnd());
            if (int e = tpc_session_set_junctions_keyed(s, (const uint64_t*)d_allj, (const uint64_t*)d_allk, total_j)) return e;
            // count pass over this GPU's position slice
            uint64_t nr = 0, ns = 0;
            if (int e = tpc_session_emit_windowed(s, slice_cut(src.n_positions, N, r), slice_cut(src.n_positions, N, r + 1), 0, 0, 0, nullptr,
                                                  nullptr, &nr, &ns)) return e;
            sh->nrec[r] = nr; sh->nstub[r] = ns;
            return 0;
        };
        rc = exchange();
        vlog("junction exchange, index, count pass done");
        if (!sh->sync_ok(r, rc)) goto done;
        uint64_t rb = 0, sb = 0;
        for (int i = 0; i < r; ++i) { rb += sh->nrec[i]; sb += sh->nstub[i]; }
        uint64_t nr = 0, ns = 0;
        rc = tpc_session_emit_windowed(s, slice_cut(src.n_positions, N, r), slice_cut(src.n_positions, N, r + 1), 1, rb, sb, window_sink, &wctx, &nr, &ns);
        if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) rc = set_error("emit failed");
        if (rc == 0) rc = tpc_session_stats(s, &sh->stats[r]);
        sh->digest[r][0] = wctx.digest[0]; sh->digest[r][1] = wctx.digest[1];
        sh->slice_bytes[r] = wctx.end;   // (end of this GPU's part of the image)
        vlog("write pass done");
        sh->sync_ok(r, rc);
        if (r == 0 && image_bytes_out) {
            uint64_t end = 0;
            for (int i = 0; i < N; ++i) end = std::max(end, sh->slice_bytes[i]);
            // a slice without records reports 0: the image ends where the last record of any slice ends
            *image_bytes_out = end;
        }
    }
done:
    cudaSetDevice(sh->dev[r]);
    if (st) cudaStreamSynchronize(st);
    if (d_allj) cudaFreeAsync(d_allj, st);
    if (d_allk) cudaFreeAsync(d_allk, st);
    if (s) { tpc_session_destroy(s); sh->session[r] = nullptr; }
    delete prov;
    if (wctx.pin) cudaFreeHost(wctx.pin);
    if (st) cudaStreamDestroy(st);
    return rc;
}

}  // namespace tpc

using namespace tpc;

struct tpc_multi {
    MultiImpl impl;
};

extern "C" {

uint32_t tpc_visible_gpus(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (uint32_t)std::max(n, 0);
}

int tpc_multi_create(uint32_t n_gpus, const int* devices, tpc_multi** out) {
    if (!out || n_gpus < 2) return set_error("a multi-GPU context needs at least 2 GPUs");
    const uint32_t vis = tpc_visible_gpus();
    if (vis == 0) return set_error("no CUDA device: twopaco_b200 has no CPU fallback");
    Nccl& nc = Nccl::get();
    if (!nc.handle) return set_error("%s", nc.error.c_str());
    tpc_multi* m = new (std::nothrow) tpc_multi();
    if (!m) return set_error("out of memory");
    for (uint32_t i = 0; i < n_gpus; ++i) {
        const int d = devices ? devices[i] : (int)i;
        if (d < 0 || (uint32_t)d >= vis) { delete m; return set_error("GPU %d is not visible (%u visible)", d, vis); }
        m->impl.dev.push_back(d);
    }
    m->impl.comm.assign(n_gpus, nullptr);
    ncclResult_t e = nc.CommInitAll(m->impl.comm.data(), (int)n_gpus, m->impl.dev.data());
    if (e != ncclSuccess) {
        m->impl.comm.assign(n_gpus, nullptr);
        delete m;
        return set_error("ncclCommInitAll failed: %s", nc.GetErrorString(e));
    }
    *out = m;
    return 0;
}

void tpc_multi_destroy(tpc_multi* m) { delete m; }

uint32_t tpc_multi_gpus(const tpc_multi* m) { return m ? (uint32_t)m->impl.dev.size() : 0; }

// position-windowed run? TPC_WINDOW_TILES=<tiles per window> forces it (0 = never); else when a GPU cannot hold the
// packed genome (0.375 B per position), the candidate and stub masks (0.25 B) and the filter side by side
static uint64_t choose_window_tiles(const tpc_multi* m, const tpc_params* params, uint64_t n_positions) {
    if (const char* e = getenv("TPC_WINDOW_TILES")) return (uint64_t)std::max(0ll, atoll(e));
    // (the size of the device's memory, asked once: cudaMemGetInfo was measured to block for tens of ms on shared hosts)
    static std::atomic<uint64_t> cached_total{0};
    uint64_t total_b = cached_total.load();
    if (!total_b) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(m->impl.dev[0]);
        size_t free_b = 0, tot = 0;
        cudaMemGetInfo(&free_b, &tot);
        cudaSetDevice(prev);
        total_b = tot;
        cached_total.store(total_b);
    }
    const double resident = 0.75 * (double)n_positions + (double)((1ull << std::max<uint32_t>(params->filter_bits, 9u)) / 8);
    return resident > 0.85 * (double)total_b ? (1u << 17) : 0;   // 2^30 positions per window
}

static int multi_run_host(tpc_multi* m, const tpc_params* params, const tpc_genome* host_genome, const ImageSink& sink,
                          uint64_t* out_bytes, tpc_stats* stats, uint64_t* digest) {
    GenomeSource src;
    src.host = host_genome;
    src.n_positions = host_genome->n_positions;
    src.rec_start = host_genome->rec_start; src.rec_len = host_genome->rec_len; src.n_records = host_genome->n_records;
    src.window_tiles = choose_window_tiles(m, params, host_genome->n_positions);
    uint64_t bytes = 0;
    int rc = run_shards(&m->impl, *params, src, sink, false, &bytes, stats, nullptr, digest);
    if (out_bytes) *out_bytes = bytes;
    if (rc == 0 && stats) stats->out_bytes = bytes;
    return rc;
}

int tpc_multi_junctions_host(tpc_multi* m, const tpc_params* params, const tpc_genome* host_genome, uint8_t* out_image,
                             uint64_t out_capacity, uint64_t* out_bytes, tpc_stats* stats) {
    if (!m || !params || !host_genome || !out_image) return set_error("null argument");
    ImageSink sink;
    sink.host_image = out_image; sink.host_capacity = out_capacity;
    uint64_t bytes = 0;
    int rc = multi_run_host(m, params, host_genome, sink, &bytes, stats, nullptr);
    if (out_bytes) *out_bytes = bytes;
    if (rc == 0 && bytes > out_capacity) {
        set_error("output buffer too small: need %llu bytes", (unsigned long long)bytes);
        return 2;
    }
    return rc;
}

int tpc_multi_junctions_digest(tpc_multi* m, const tpc_params* params, const tpc_genome* host_genome, uint64_t digest[2],
                               uint64_t* image_bytes, tpc_stats* stats) {
    if (!m || !params || !host_genome || !digest) return set_error("null argument");
    ImageSink sink;
    sink.digest_only = true;
    return multi_run_host(m, params, host_genome, sink, image_bytes, stats, digest);
}

}  // extern "C"

// internal entry point of tpc_build: the packed genome sits on GPU m->dev[0]; the image goes to `fd`
int tpc::multi_run_from_device0(tpc_multi* m, const tpc_params* params, const uint64_t* dev0_codes, const uint64_t* dev0_nmask,
                                uint64_t n_positions, const uint64_t* rec_start, const uint64_t* rec_len, uint64_t n_records, int fd,
                                uint64_t* image_bytes, tpc_stats* stats, tpc_session** session0) {
    GenomeSource src;
    src.dev0_codes = dev0_codes; src.dev0_nmask = dev0_nmask;
    src.n_positions = n_positions; src.rec_start = rec_start; src.rec_len = rec_len; src.n_records = n_records;
    ImageSink sink;
    sink.fd = fd;
    return run_shards(&m->impl, *params, src, sink, true, image_bytes, stats, session0);
}
