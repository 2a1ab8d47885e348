// tpc_windowed.inl -- position-windowed sessions (included at the end of tpc_session.cu): inputs whose packed genome,
// candidate mask and ownership planes do not fit HBM beside the filter (BASELINE config 5: 310 Gbp, a 128 GiB filter per
// GPU).  The reference handles any input size by re-reading the FASTA files chunk by chunk in every stage
// (DistributeTasks, vertexenumerator.h:1108-1226, once per stage and round h:228-392); here the PACKED genome streams
// through HBM in windows of whole tiles, once per pass, from a WindowProvider (pinned host memory over PCIe; several
// GPUs: 1/N per GPU + NCCL all-gather), while the filter, the candidate table and the junction index stay resident:
//
//   per round (hash part):  fill pass   : for every window  k_fill
//                           query pass  : for every window  k_query -> window-local mask -> k_insert into the round's table
//                           k_classify  -> junction (first position, key) pairs
//   index                :  from the keys (no genome access: k <= 31, the canonical k-mer sits in the slot)
//   emit (position slice):  for every window  every definite k-mer position is tested against the index (no candidate
//                           mask is kept: it would be one bit per position) -> the window's records in image order
//
// Window-local arrays are addressed through VIRTUAL base pointers (buffer - first word of the window), so the kernels
// run unchanged on absolute tile / word / position indices.  Restrictions: k <= 31 and no -a (the table and index slots
// then hold the k-mer itself; the position-identified slots of longer k-mers read the genome back at arbitrary
// positions), the direct filter kernels (the binned path keeps ownership planes and records of the whole input).
namespace tpc {

static uint32_t floor_log2_u64(uint64_t x) {
    uint32_t l = 0;
    while ((2ull << l) <= x) ++l;
    return l;
}

// point the session's genome / mask views at a window's buffers
static void windowed_view(tpc_session* s, const uint64_t* codes_v, const uint64_t* nmask_v, uint64_t tile_begin) {
    s->g = GenomeView{codes_v, nmask_v, s->w_npos};
    s->d_mask = s->d_wmask - tile_begin * kTileThreads;
    s->d_stubmask = s->d_wstub - tile_begin * kTileThreads;
}

static int windowed_buffers(tpc_session* s) {
    const uint64_t words = s->window_tiles * kTileThreads + 64;
    if (!s->d_wmask) CK(dev_alloc(&s->d_wmask, words * 4, s->stream));
    if (!s->d_wstub) CK(dev_alloc(&s->d_wstub, words * 4, s->stream));
    return 0;
}

static int find_candidates_windowed(tpc_session* s) {
    if (!s->inline_keys())
        return set_error("a windowed run (input larger than device memory) needs k <= 31 and no abundance limit");
    WallTimer wall(&s->st.ms_wall_candidates);
    LaunchCtx lc = s->lctx();
    WindowProvider* wp = s->wp;
    const uint64_t filter_bytes = (1ull << s->filter_bits_eff) / 8;
    if (!s->d_filter) CK(dev_alloc(&s->d_filter, filter_bytes, s->stream));
    if (int rc = windowed_buffers(s)) return rc;
    CK(cudaMemsetAsync(s->d_ctr, 0, sizeof(Counters), s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->local_count = 0;
    s->sub_rounds = 1;
    s->rounds_eff = s->prm.rounds;
    s->st.sub_rounds = 1;
    s->own = OwnPlanes{};
    s->own_shared = false;
    // the round's candidate table: sized from the memory left beside the filter (the marks are only known after the
    // query pass, which inserts as it goes); a table that turns out too small doubles and the round is redone
    uint32_t lg = std::min<uint32_t>(floor_log2_u64(std::max<uint64_t>(available_bytes(s->device) / 4 / sizeof(Slot), 1024)), 31);
    // (never more than two slots per position of a round's share of the input)
    lg = std::min(lg, std::max<uint32_t>(10, ceil_log2(2 * s->w_npos / ((uint64_t)s->prm.rounds * s->prm.shard_count) + 16)));
    if (const char* e = getenv("TPC_WINDOW_TABLE_LOG2")) lg = (uint32_t)std::min(31, std::max(4, atoi(e)));   // (tests: the redo path)
    float ms_fill = 0, ms_query = 0, ms_classify = 0;
    Counters prev{}, cur{};
    Events evs(4);
    for (uint32_t r = 0; r < s->rounds_eff; ++r) {
        KParams kp = s->kparams(s->prm.shard_index * s->rounds_eff + r);
        for (;;) {   // (redone with a larger table when the insert overflowed)
            const uint64_t need = sizeof(Slot) << lg;
            if (s->d_T && need != s->T_bytes) { CK(dev_free(s->d_T, s->stream)); s->d_T = nullptr; }
            if (!s->d_T) {
                if (need > available_bytes(s->device)) return set_error("candidate table of 2^%u slots does not fit in device memory: use more rounds (-r)", lg);
                CK(dev_alloc(&s->d_T, need, s->stream));
                s->T_bytes = need;
            }
            s->T_log2 = lg;
            const TableView T{s->d_T, lg, 1u};
            CK(cudaMemsetAsync(s->d_filter, 0, filter_bytes, s->stream));   // h:257: zero-filled each round
            CK(cudaMemsetAsync(s->d_T, 0, need, s->stream));
            for (int pass = 0; pass < 2; ++pass) {
                CK(cudaEventRecord(evs[pass], s->stream));
                wp->begin_pass(WindowProvider::kShared, 0, s->ntiles);
                for (uint64_t t0 = 0; t0 < s->ntiles; t0 += s->window_tiles) {
                    const uint64_t t1 = std::min(s->ntiles, t0 + s->window_tiles);
                    const uint64_t *cv = nullptr, *nv = nullptr;
                    cudaEvent_t ready = nullptr;
                    if (int rc = wp->fetch(t0, t1, &cv, &nv, &ready)) return rc;
                    if (ready) CK(cudaStreamWaitEvent(s->stream, ready, 0));
                    windowed_view(s, cv, nv, t0);
                    if (pass == 0) {
                        CK(W_DISPATCH(s, fill(lc, s->g, s->d_filter, kp, t0, t1, s->d_ctr)));
                    } else {
                        // (the query kernel skips the words past the last position: they must not hold an older window's marks)
                        CK(cudaMemsetAsync(s->d_wmask, 0, (s->window_tiles * kTileThreads + 64) * 4, s->stream));
                        CK(W_DISPATCH(s, query(lc, s->g, s->d_filter, kp, t0, t1, s->d_mask, 0, s->d_ctr, s->d_hll)));
                        KParams kpi = kp;
                        kpi.nparts = 1;   // the window mask holds this round's marks only
                        CK(W_DISPATCH(s, insert(lc, s->g, s->d_mask, kpi, t0, t1, T, s->d_ctr, nullptr)));
                    }
                    CK(cudaStreamSynchronize(s->stream));   // the window's buffers may be reused
                    wp->release(t0);
                }
            }
            CK(cudaEventRecord(evs[2], s->stream));
            CK(cudaMemcpyAsync(&cur, s->d_ctr, sizeof cur, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaStreamSynchronize(s->stream));
            float t = 0;
            cudaEventElapsedTime(&t, evs[0], evs[1]); ms_fill += t;
            cudaEventElapsedTime(&t, evs[1], evs[2]); ms_query += t;
            const bool too_small = !(cur.overflow == prev.overflow && (cur.distinct - prev.distinct) * 10 <= (7ull << lg));
            if (!wp->any_shard(too_small)) break;   // (shards share the passes: they repeat a round together)
            Counters redo = prev;   // forget the round: marks, filter statistics and table counters
            CK(cudaMemcpyAsync(s->d_ctr, &redo, sizeof redo, cudaMemcpyHostToDevice, s->stream));
            cur = redo;
            if (too_small) {
                if (lg >= 31) return set_error("too many candidates for one round: use more rounds (-r)");
                ++lg;
            }
        }
        const TableView T{s->d_T, s->T_log2, 1u};
        const uint64_t distinct_r = cur.distinct - prev.distinct;
        if (s->local_count + distinct_r > s->local_cap) {
            const uint64_t ncap = std::max<uint64_t>(s->local_count + distinct_r * (s->rounds_eff - r) * 9 / 8, 1024);
            unsigned long long *nl = nullptr, *nk = nullptr;
            CK(dev_alloc(&nl, ncap * 8, s->stream));
            CK(dev_alloc(&nk, ncap * 8, s->stream));
            if (s->local_count) {
                CK(cudaMemcpyAsync(nl, s->d_local, s->local_count * 8, cudaMemcpyDeviceToDevice, s->stream));
                CK(cudaMemcpyAsync(nk, s->d_local_keys, s->local_count * 8, cudaMemcpyDeviceToDevice, s->stream));
            }
            CK(cudaStreamSynchronize(s->stream));
            if (s->d_local) CK(dev_free(s->d_local, s->stream));
            if (s->d_local_keys) CK(dev_free(s->d_local_keys, s->stream));
            s->d_local = nl; s->d_local_keys = nk; s->local_cap = ncap;
        }
        CK(cudaEventRecord(evs[2], s->stream));
        CK(launch_classify(lc, T, s->prm.abundance, 0, s->d_local, s->d_local_keys, s->local_cap, s->d_ctr));
        CK(cudaEventRecord(evs[3], s->stream));
        CK(cudaMemcpyAsync(&cur, s->d_ctr, sizeof cur, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        float t = 0;
        cudaEventElapsedTime(&t, evs[2], evs[3]); ms_classify += t;
        s->local_count = cur.junctions;
        prev = cur;
    }
    if (s->d_T) { CK(dev_free(s->d_T, s->stream)); s->d_T = nullptr; s->T_bytes = 0; }
    // the filter is not needed any more; the junction index wants the memory
    CK(dev_free(s->d_filter, s->stream));
    s->d_filter = nullptr;
    s->st.candidate_marks = cur.marks;
    s->st.candidate_kmers = cur.distinct;
    s->st.filter_edges_set = cur.filter_new;
    s->st.ms_fill = ms_fill; s->st.ms_query = ms_query; s->st.ms_classify = ms_classify;   // (ms_query includes the inserts)
    s->have_candidates = true;
    s->have_index = false;
    return 0;
}

}  // namespace tpc

extern "C" {

int tpc_session_set_genome_windowed(tpc_session* s, uint64_t n_positions, const uint64_t* rec_start, const uint64_t* rec_len,
                                    uint64_t n_records, uint64_t window_tiles, tpc::WindowProvider* provider) {
    if (!s || !provider || (n_records && (!rec_start || !rec_len))) return set_error("null argument");
    if (s->g.codes || s->windowed) return set_error("genome already set");
    if (window_tiles == 0) return set_error("a window holds at least one tile");
    tpc_genome g{};
    g.n_positions = n_positions; g.rec_start = rec_start; g.rec_len = rec_len; g.n_records = n_records;
    if (int rc = adopt_records(s, &g)) return rc;
    s->windowed = true;
    s->wp = provider;
    s->window_tiles = window_tiles;
    s->w_npos = n_positions;
    return 0;
}

int tpc_session_local_junction_keys(tpc_session* s, const uint64_t** dev_keys) {
    if (!s || !s->have_candidates) return set_error("find_candidates has not run");
    if (dev_keys) *dev_keys = (const uint64_t*)s->d_local_keys;   // null unless the run is windowed
    return 0;
}

// BifurcationStorage::Init from (first position, key) pairs of ALL shards: ids = rank of the first position, the
// index is built from the keys alone (windowed runs).
int tpc_session_set_junctions_keyed(tpc_session* s, const uint64_t* dev_words_all, const uint64_t* dev_keys_all, uint64_t n) {
    if (!s || !s->windowed) return set_error("not a windowed session");
    if (n && (!dev_words_all || !dev_keys_all)) return set_error("null argument");
    LaunchCtx lc = s->lctx();
    WallTimer wall(&s->st.ms_wall_index);
    CK(cudaEventRecord(s->ev[5], s->stream));
    if (s->d_sorted) { CK(dev_free(s->d_sorted, s->stream)); s->d_sorted = nullptr; }
    if (s->d_J) { CK(dev_free(s->d_J, s->stream)); s->d_J = nullptr; }
    CK(dev_alloc(&s->d_sorted, std::max<uint64_t>(n, 1) * 8, s->stream));
    unsigned long long* keys_sorted = nullptr;
    CK(dev_alloc(&keys_sorted, std::max<uint64_t>(n, 1) * 8, s->stream));
    DevBuf holder;
    holder.p = keys_sorted; holder.st = s->stream;
    if (n) {
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned long long*)dev_words_all, s->d_sorted,
                                           (const unsigned long long*)dev_keys_all, keys_sorted, n, 0, kPosBits, s->stream));
        if (tmp > s->sort_tmp_bytes) {
            if (s->d_sort_tmp) CK(dev_free(s->d_sort_tmp, s->stream));
            CK(dev_alloc(&s->d_sort_tmp, tmp, s->stream));
            s->sort_tmp_bytes = tmp;
        }
        CK(cub::DeviceRadixSort::SortPairs(s->d_sort_tmp, tmp, (const unsigned long long*)dev_words_all, s->d_sorted,
                                           (const unsigned long long*)dev_keys_all, keys_sorted, n, 0, kPosBits, s->stream));
    }
    s->J_log2 = std::max<uint32_t>(ceil_log2(n * 2 + 16), 6);
    CK(dev_alloc(&s->d_J, sizeof(Slot) << s->J_log2, s->stream));
    CK(cudaMemsetAsync(s->d_J, 0, sizeof(Slot) << s->J_log2, s->stream));
    if (n) {
        const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)s->sm_count * 8);
        k_build_index_keys<<<grid, 256, 0, s->stream>>>(keys_sorted, n, s->kparams(0), TableView{s->d_J, s->J_log2, 1u});
        ++s->launches;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(s->ev[6], s->stream));
    CK(cudaStreamSynchronize(s->stream));   // keys_sorted is released on return
    s->J_count = n;
    s->st.junctions = n;
    s->have_index = true;
    s->have_count = false;
    return 0;
}

// EdgeConstructionWorker + JunctionPositionWriter over the positions [pos_begin, pos_end) of a windowed session, window by
// window.  write == 0: count only (*n_records / *n_stubs: totals of the slice).  write != 0: every window's part of the
// image is produced in a device buffer and handed to `sink(ctx, dev_bytes, image_offset, nbytes)` in image order
// (records_before / stubs_before: totals of the slices before this one).  The provider is asked for kPrivate windows:
// only this GPU reads them.
int tpc_session_emit_windowed(tpc_session* s, uint64_t pos_begin, uint64_t pos_end, int write, uint64_t records_before,
                              uint64_t stubs_before, tpc_window_sink sink, void* ctx, uint64_t* n_records, uint64_t* n_stubs) {
    if (!s || !s->windowed || !s->have_index) return set_error("set_junctions_keyed has not run");
    if (pos_end > s->w_npos) pos_end = s->w_npos;
    if (pos_begin > pos_end) pos_begin = pos_end;
    if (pos_begin % kTilePos && pos_begin != s->w_npos) return set_error("emit slice must start at a multiple of %d", kTilePos);
    if (int rc = windowed_buffers(s)) return rc;
    LaunchCtx lc = s->lctx();
    const KParams kp = s->kparams(0);
    const uint64_t tb = pos_begin / kTilePos, te = pos_begin >= pos_end ? tb : (pos_end + kTilePos - 1) / kTilePos;
    uint64_t rb = records_before, sb = stubs_before, tot_rec = 0, tot_stub = 0;
    uint8_t* d_out = nullptr;
    uint64_t out_cap = 0;
    DevBuf out_holder;
    out_holder.st = s->stream;
    s->wp->begin_pass(WindowProvider::kPrivate, tb, te);
    for (uint64_t t0 = tb; t0 < te; t0 += s->window_tiles) {
        const uint64_t t1 = std::min(te, t0 + s->window_tiles);
        const uint64_t *cv = nullptr, *nv = nullptr;
        cudaEvent_t ready = nullptr;
        if (int rc = s->wp->fetch(t0, t1, &cv, &nv, &ready)) return rc;
        if (ready) CK(cudaStreamWaitEvent(s->stream, ready, 0));
        windowed_view(s, cv, nv, t0);
        // no candidate mask is kept in a windowed run: every definite k-mer position is resolved against the index
        CK(W_DISPATCH(s, valid_mask(lc, s->g, kp, t0, t1, s->d_mask)));
        uint64_t nr = 0, ns = 0;
        const uint64_t p0 = std::max(pos_begin, t0 * kTilePos), p1 = std::min(pos_end, t1 * kTilePos);
        if (int rc = tpc_session_emit_count(s, p0, p1, &nr, &ns)) return rc;
        if (write) {
            const uint64_t cap = 12 * (nr + s->rec_start.size()) + 16;
            if (cap > out_cap) {
                if (d_out) CK(dev_free(d_out, s->stream));
                d_out = nullptr; out_holder.p = nullptr;
                CK(dev_alloc(&d_out, cap + cap / 4, s->stream));
                out_holder.p = d_out;
                out_cap = cap + cap / 4;
            }
            uint64_t off = 0, nb = 0;
            if (int rc = tpc_session_emit_write(s, rb, sb, d_out, out_cap, &off, &nb)) return rc;
            if (sink)
                if (int rc = sink(ctx, d_out, off, nb, s->stream)) return rc;
        }
        CK(cudaStreamSynchronize(s->stream));
        s->wp->release(t0);
        rb += nr; sb += ns; tot_rec += nr; tot_stub += ns;
    }
    if (n_records) *n_records = tot_rec;
    if (n_stubs) *n_stubs = tot_stub;
    s->st.occurrences = tot_rec;
    s->st.stubs = tot_stub;
    return 0;
}

}  // extern "C"
