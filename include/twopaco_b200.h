/*
 * twopaco_b200.h -- C ABI of libtwopaco_b200.so: TwoPaCo's two-pass junction-finding hot
 * path on NVIDIA B200 (sm_100a).  Plain pointers and sizes only; no C++/torch types cross
 * this boundary; nothing throws across it (errors: non-zero return + tpc_last_error()).
 *
 * Citations are relative to the reference tree (medvedevgroup/TwoPaCo, /root/reference).
 *
 * Three levels, top to bottom:
 *   1. tpc_build()            == TwoPaCo::CreateEnumerator(...)   (FASTA files -> de_bruijn.bin)
 *   2. tpc_junctions_host()   == the same work on an already packed genome in host memory
 *                                (the "e2e" timed region of bench.py: H2D + kernels + D2H)
 *   3. tpc_session_*()        == the stages on a device-resident genome, one session per GPU
 *                                (hash-range shard), so that the caller can put collectives
 *                                between the stages (multi-GPU) and time kernels alone.
 */
#ifndef TWOPACO_B200_H_
#define TWOPACO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPC_ABI_VERSION 4
#define TPC_INVALID_VERTEX INT64_MAX          /* src/graphconstructor/common.cpp:5 */
#define TPC_SEPARATOR_POS 0xFFFFFFFFu         /* src/common/junctionapi.h:36-37 */
#define TPC_MAX_K 603                         /* up to 19 x 64-bit words per packed k-mer: the reference's limit
                                                 (MAX_CAPACITY 20, vertexenumerator.h:4; capacity = ceil((k + 4) / 32)
                                                 <= 19, candidateoccurence.h:129-133; larger k: "K is too big") */
#define TPC_STUB_ID_OFFSET 42                 /* vertexenumerator.h:419 */

typedef struct tpc_handle tpc_handle;         /* result of tpc_build (a VertexEnumerator) */
typedef struct tpc_session tpc_session;       /* one GPU's shard of a run */
typedef struct tpc_multi tpc_multi;           /* the GPUs of this process + their NCCL communicators */
typedef void (*tpc_log_fn)(void *ctx, const char *text); /* receives the std::ostream & log text */

/* ------------------------------------------------------------------------------------------
 * Parameters of one run == the arguments of CreateEnumerator
 * (src/graphconstructor/vertexenumerator.h:37-46; CLI: constructor.cpp:60-143).
 * ---------------------------------------------------------------------------------------- */
typedef struct tpc_params {
    uint32_t k;            /* vertexLength, odd, 1..TPC_MAX_K            (-k, default 25)      */
    uint32_t filter_bits;  /* filterSize: the Bloom filter has 2^f bits   (-f)                  */
    uint32_t q;            /* hashFunctions: bits set per edge            (-q, default 5)       */
    uint32_t rounds;       /* rounds: hash-range rounds run in sequence   (-r, default 1)       */
    uint64_t abundance;    /* junctions seen more often are dropped       (-a, default 2^64-1)  */
    uint32_t shard_index;  /* this GPU's hash-range shard, 0..shard_count-1 (spatial analogue   */
    uint32_t shard_count;  /*   of -r: vertexenumerator.h:234-254, 638, 1066); 1 = unsharded    */
    uint64_t seed;         /* hash seed; fixed default 0 (the reference seeds from /dev/urandom, */
                           /*   mersennetwister.h:242-263, which only permutes ids)             */
} tpc_params;

/* Counters the reference prints to its log (vertexenumerator.h:384-387, 463) plus timings. */
typedef struct tpc_stats {
    uint64_t positions;          /* genome positions incl. separators                         */
    uint64_t candidate_marks;    /* "Candidate marks count"                                   */
    uint64_t candidate_kmers;    /* "Hash table size" (distinct candidate k-mers)             */
    uint64_t junctions;          /* "True junctions count" == "Distinct junctions"            */
    uint64_t occurrences;        /* "True marks count" (records written, incl. stubs)         */
    uint64_t stubs;              /* records that are end-of-sequence stubs                    */
    uint64_t out_bytes;          /* size of the de_bruijn.bin image                           */
    uint64_t filter_edges_set;   /* fill: (vertex, edge-slot) items that set a new bit        */
    float ms_bin, ms_fill, ms_query, ms_insert, ms_classify, ms_index, ms_emit, ms_total; /* CUDA events;
                                    ms_bin = partition of filter records by slice (binned path only) */
    uint32_t kernel_launches;    /* kernels of this library launched by the call              */
    uint32_t bin_waves;          /* binned path: waves of records per pass (1 = records shared by
                                    fill and query); 0 = direct path                             */
    uint32_t sub_rounds;         /* hash sub-ranges per -r round chosen so that one round's records
                                    fit HBM in one wave (1 = none); unobservable in the output    */
    float ms_bin_overlapped;     /* pipelined rounds: binning of round r+1 that ran beside the fill of round r
                                    (not part of ms_bin / ms_total) */
    float ms_wall_candidates;    /* host wall clock of tpc_session_find_candidates (kernels + memsets + allocations +
                                    host synchronisations), so that time outside the CUDA-event stage times is visible */
    float ms_wall_index, ms_wall_emit;   /* the same for set_junctions and emit_count + emit_write */
    uint32_t skew_rebins;        /* binned path: rounds whose slices overflowed (repeat-rich input) and were re-binned into
                                    arrays sized from the exact per-slice record counts */
    uint64_t h2d_bytes;          /* host-buffer entry points: genome bytes actually copied host -> device (uniform blocks of
                                    the n-mask are set on the device instead of copied) */
} tpc_stats;

/* ------------------------------------------------------------------------------------------
 * Packed genome (2-bit DNA packing of src/graphconstructor/compressedstring.h:239-264: base i
 * -> bits 2(i%32) of 64-bit word i/32, A0 C1 G2 T3, dnachar.cpp:18-33), applied to the WHOLE
 * input instead of per k-mer:
 *   all records are laid out in one position space, separated by one 'N' position -- the
 *   sentinel 'N' the reference puts before and after every record (vertexenumerator.h:1154,
 *   1191):   pos 0 = N, record 0 at [1, 1+len0), N, record 1, N, ...
 *   codes : 2 bits / position (0 where N), n_mask : 1 bit / position (1 = not ACGT).
 * Both arrays must be readable up to the word counts returned by the helpers below.
 * ---------------------------------------------------------------------------------------- */
typedef struct tpc_genome {
    const uint64_t *codes;      /* tpc_code_words(n_positions) words                            */
    const uint64_t *n_mask;     /* tpc_mask_words(n_positions) words                            */
    uint64_t n_positions;       /* 1 + sum(len_i + 1)                                           */
    const uint64_t *rec_start;  /* n_records position of the first base of each record          */
    const uint64_t *rec_len;    /* n_records lengths (< 2^32: junctionapi.h:33-34)              */
    uint64_t n_records;
} tpc_genome;

uint64_t tpc_code_words(uint64_t n_positions);  /* incl. the read-ahead padding the kernels need */
uint64_t tpc_mask_words(uint64_t n_positions);
uint64_t tpc_positions_for(const uint64_t *rec_len, uint64_t n_records);

/* Host packer: normalised records (bytes over ACGTN, anything else -> N; lower case folded,
 * DistributeTasks vertexenumerator.h:1174) -> codes / n_mask / rec_start at the layout above.
 * Outputs are caller-allocated (sizes from the helpers above); threads = host worker threads. */
int tpc_pack_records(const char *const *records, const uint64_t *rec_len, uint64_t n_records,
                     uint32_t threads, uint64_t *codes, uint64_t *n_mask, uint64_t *rec_start);

/* FASTA framing of src/common/streamfastaparser.cpp:29-133 (header = line starting with '>',
 * whitespace skipped, upper-cased, characters outside ACGTURYKMSWBDHWNXV are an error).
 * Appends the records of one file; *records / *rec_len are malloc'ed/realloc'ed by the callee
 * and released with tpc_free_records. */
int tpc_read_fasta(const char *path, char ***records, uint64_t **rec_len, uint64_t *n_records);
void tpc_free_records(char **records, uint64_t *rec_len, uint64_t n_records);

/* Multi-threaded whole-input parser used by tpc_build: all records of all files, normalised, in
 * ONE buffer in the position layout of tpc_genome (1 byte per position, 'N' separators; feed it to
 * tpc_pack_ascii_device).  Outputs are malloc'ed by the callee; release with tpc_host_free. */
int tpc_ingest_fasta(const char *const *paths, size_t n_files, uint32_t threads, uint8_t **ascii,
                     uint64_t *n_positions, uint64_t **rec_start, uint64_t **rec_len,
                     uint64_t *n_records);
void tpc_host_free(void *p);

/* ------------------------------------------------------------------------------------------
 * Level 1 -- replaces  std::unique_ptr<VertexEnumerator> TwoPaCo::CreateEnumerator(fileName,
 * vertexLength, filterSize, hashFunctions, rounds, threads, abundance, tmpDirName,
 * outFileName, logStream)   (vertexenumerator.h:37-46, vertexenumerator.cpp:73-94).
 * Blocking; all work happens inside the call (as in VertexEnumeratorImpl's constructor,
 * vertexenumerator.h:122-466); writes `outfile` in the de_bruijn.bin format
 * (junctionapi.h:107-137).  `threads` = host threads for FASTA parsing/packing; the GPU is
 * the current CUDA device.  `tmpdir` is accepted for CLI compatibility (no temp files are
 * needed: candidate masks and junction keys stay in HBM; reference: h:219, 292).
 * Error strings follow the reference ("Can't open file ...", "Found an invalid character",
 * "The value of K is too big. ...", "Can't create the output file").
 * ---------------------------------------------------------------------------------------- */
int tpc_build(const char *const *fasta_paths, size_t n_files, uint32_t k, uint32_t filter_bits,
              uint32_t q, uint32_t rounds, uint32_t threads, uint64_t abundance,
              const char *tmpdir, const char *outfile, tpc_log_fn log, void *log_ctx,
              tpc_handle **out);
uint64_t tpc_vertices(const tpc_handle *h);                 /* VertexEnumerator::GetVerticesCount, h:104-107 */
int64_t tpc_get_id(const tpc_handle *h, const char *kmer);  /* VertexEnumerator::GetId, h:98-102;
                                                               TPC_INVALID_VERTEX when absent           */
int tpc_handle_stats(const tpc_handle *h, tpc_stats *out);
void tpc_free(tpc_handle *h);

/* ------------------------------------------------------------------------------------------
 * Level 2 -- packed genome in HOST memory -> de_bruijn.bin image in HOST memory, one GPU.
 * out_image (capacity bytes; pinned memory recommended) receives the exact file content.
 * Returns 0, or 2 when out_capacity is too small (*out_bytes then holds the required size).
 * ---------------------------------------------------------------------------------------- */
int tpc_junctions_host(const tpc_params *params, const tpc_genome *host_genome,
                       uint8_t *out_image, uint64_t out_capacity, uint64_t *out_bytes,
                       tpc_stats *stats);

/* The same on N GPUs of this process (hash-range shards, one host thread and one session per GPU):
 * every GPU uploads 1/N of the packed genome over its own PCIe link, the parts are all-gathered chunk by chunk over
 * NVLink (NCCL) while the first pass already runs, junction lists are all-gathered, candidate masks OR-reduce-
 * scattered, and every GPU copies its position slice of the image into out_image.  A context is created once
 * (ncclCommInitAll over `devices`, NULL = 0..n_gpus-1) and reused by any number of runs.  This is what tpc_build
 * uses when several GPUs are visible.  NCCL is loaded at run time (libnccl.so.2; TPC_NCCL_LIB overrides). */
uint32_t tpc_visible_gpus(void);
int tpc_multi_create(uint32_t n_gpus, const int *devices, tpc_multi **out);
void tpc_multi_destroy(tpc_multi *m);
uint32_t tpc_multi_gpus(const tpc_multi *m);
int tpc_multi_junctions_host(tpc_multi *m, const tpc_params *params, const tpc_genome *host_genome,
                             uint8_t *out_image, uint64_t out_capacity, uint64_t *out_bytes, tpc_stats *stats);
/* The same run whose image never leaves the GPUs: only its size and its position-keyed digest (tpc_image_digest_device)
 * come back -- for outputs too large to keep (BASELINE config 5 writes ~0.7 TB of records) and for timing the path
 * without the device-to-host copy. */
int tpc_multi_junctions_digest(tpc_multi *m, const tpc_params *params, const tpc_genome *host_genome, uint64_t digest[2],
                               uint64_t *image_bytes, tpc_stats *stats);
/* Inputs that do not fit a GPU's memory beside the filter (packed genome 0.375 B + masks 0.25 B per position) run
 * position-windowed, in tpc_junctions_host and tpc_multi_*: the packed genome stays in (pinned) host memory and streams
 * through HBM window by window, once per pass (fill, query + exact insert per round; emit), the filter, the candidate
 * table and the junction index stay resident (reference: DistributeTasks re-reads the FASTA once per stage,
 * vertexenumerator.h:1108-1226).  Several GPUs: every GPU uploads 1/N of a window over its own PCIe link, NCCL
 * all-gathers it.  Needs k <= 31 and no -a.  TPC_WINDOW_TILES=<8192-position tiles per window> forces it (0: never). */

/* ------------------------------------------------------------------------------------------
 * Level 3 -- sessions.  One session = one GPU (current device at creation) = one hash-range
 * shard (params->shard_index / shard_count).  All device pointers returned stay owned by the
 * session.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *
 *   create -> set_genome_{host,device} -> find_candidates -> local_junctions
 *          [multi-GPU: all-gather the junction lists]      -> set_junctions
 *          [multi-GPU: OR-reduce the candidate masks]      -> emit_count -> emit_write
 * ---------------------------------------------------------------------------------------- */
int tpc_session_create(const tpc_params *params, void *stream, tpc_session **out);
void tpc_session_destroy(tpc_session *s);

/* Copy a packed genome host->device (async on the session stream; buffers must stay alive
 * until the next synchronising call) or adopt device-resident arrays (not copied, not freed;
 * rec_start / rec_len are HOST arrays in both cases). */
int tpc_session_set_genome_host(tpc_session *s, const tpc_genome *host_genome);
int tpc_session_set_genome_device(tpc_session *s, const tpc_genome *genome_with_device_arrays);

/* A device-resident genome handed to tpc_session_set_genome_device may still be arriving (multi-GPU:
 * every GPU uploads 1/N of it and the parts are all-gathered chunk by chunk over NVLink).  Declare
 * its chunks: tiles (8192 positions) from tile_begin up to the next chunk's tile_begin are complete
 * once `cuda_event` (a cudaEvent_t, recorded by the producer on its stream) has completed.  Chunks are
 * added in ascending order, the first at tile 0, before tpc_session_find_candidates; the first pass
 * over the genome then starts on the chunks that have arrived.  The events stay owned by the caller
 * and must outlive the session's use of the genome. */
int tpc_session_add_genome_event(tpc_session *s, uint64_t tile_begin, void *cuda_event);

/* Pass 1 + pass 2 for this shard (reference stages 1a, 1b, 2: FilterFillerWorker h:995-1105,
 * CandidateCheckingWorker h:586-704, CandidateFinalFilteringWorker h:708-829), `rounds` times
 * over disjoint hash sub-ranges. */
int tpc_session_find_candidates(tpc_session *s);

/* TrueBifurcations (h:1228-1256): this shard's junctions as a device array of 64-bit words
 * (low 40 bits: position of the first occurrence of the k-mer in the genome; the k-mer is
 * read back from the genome, so the word is meaningful on every GPU holding the genome). */
int tpc_session_local_junctions(tpc_session *s, const uint64_t **dev_words, uint64_t *count);

/* BifurcationStorage::Init (bifurcationstorage.h:27-66): build the id index from the junction
 * words of ALL shards (device array).  Ids are 1..J in order of first occurrence. */
int tpc_session_set_junctions(tpc_session *s, const uint64_t *dev_words_all, uint64_t count_all);

/* Candidate mask of this shard: 1 bit per position, 32-bit words (bit i of word w = position
 * 32w+i).  Masks of different shards are disjoint; OR (== sum) them before emitting.  n_words is
 * padded (with zero words) to shard_count equal chunks of whole 8192-position tiles, chunk r being
 * the slice shard r emits: a reduce-scatter of the shards' masks lands each chunk where
 * tpc_session_emit_count reads it (only the words of the emitted slice have to be complete). */
int tpc_session_candidate_mask(tpc_session *s, uint32_t **dev_mask, uint64_t *n_words);

/* EdgeConstructionWorker (h:856-993) + JunctionPositionWriter (junctionapi.h:107-137) for the
 * positions [pos_begin, pos_end) (multiples of 8192, or n_positions).  emit_count returns the
 * number of junction records and stubs of the slice; emit_write writes the slice's part of the
 * file image (records + the separators that precede its records) into dev_out, given how many
 * records / stubs precede the slice.  Image offset of the slice = 12 * (records_before +
 * index of the sequence of its first record ... ) is returned in *image_offset. */
int tpc_session_emit_count(tpc_session *s, uint64_t pos_begin, uint64_t pos_end,
                           uint64_t *n_records, uint64_t *n_stubs);
int tpc_session_emit_write(tpc_session *s, uint64_t records_before, uint64_t stubs_before,
                           uint8_t *dev_out, uint64_t out_capacity,
                           uint64_t *image_offset, uint64_t *image_bytes);

int tpc_session_get_id(tpc_session *s, const char *kmer, int64_t *id);
int tpc_session_stats(tpc_session *s, tpc_stats *out);   /* synchronises the stream */

/* ------------------------------------------------------------------------------------------
 * K0: ASCII -> 2-bit codes + N mask on the device.  dev_ascii holds one byte per position in
 * the layout of tpc_genome ('N' or any non-ACGT byte at separators; case folded; the alphabet
 * of dnachar.cpp:18-33).  The reference has no such step: it re-parses the ASCII FASTA once per
 * stage (vertexenumerator.h:1135-1214).
 * ---------------------------------------------------------------------------------------- */
int tpc_pack_ascii_device(const uint8_t *dev_ascii, uint64_t n_positions, uint64_t *dev_codes,
                          uint64_t *dev_nmask, void *stream);

/* Position-keyed digest of (a slice of) a de_bruijn.bin image in device memory: two 64-bit sums over the
 * image's 32-bit words of mix(global word index, word).  The digests of disjoint slices of one image add up
 * (mod 2^64) to the digest of the whole image, so N GPUs that each hold a slice (tpc_session_emit_write:
 * image_offset / image_bytes) can prove byte-identity with a single-GPU run without gathering the image.
 * nbytes and image_offset are multiples of 4.  Synchronises `stream`. */
int tpc_image_digest_device(const uint8_t *dev_image, uint64_t nbytes, uint64_t image_offset, void *stream,
                            uint64_t digest[2]);

/* ------------------------------------------------------------------------------------------
 * The consumer side of de_bruijn.bin on the GPU.
 *   tpc_graphdump_device : the text `graphdump -f seq` (format 0; graphdump.cpp:160-168: "chr pos id" per record) or
 *                          `graphdump -f group` (format 1; graphdump.cpp:120-158: occurrences grouped by signed id,
 *                          "chr pos; " per member, classes ordered by first occurrence) or `graphdump -f dot` (format 2;
 *                          graphdump.cpp:585-606: two edge lines per consecutive pair of records) prints for an image in device
 *                          memory; *dev_text is allocated by the callee (release with tpc_device_free).
 *   tpc_graphdump_file   : file -> text file (out_path NULL or "-" = stdout), format "seq" | "group" | "dot".
 *   tpc_canonical_image_device : the canonical relabelling of SURVEY.md appendix C -- ids renumbered 1.. by first
 *                          appearance of |id|, first occurrence positive, separators kept -- written to dev_out (same size).
 *                          Two images describe the same graph iff their canonical images are byte-identical.
 * ---------------------------------------------------------------------------------------- */
int tpc_graphdump_device(const uint8_t *dev_image, uint64_t image_bytes, uint32_t format, void *stream,
                         uint8_t **dev_text, uint64_t *text_bytes);
int tpc_graphdump_file(const char *image_path, const char *format, const char *out_path);
int tpc_canonical_image_device(const uint8_t *dev_image, uint64_t image_bytes, void *stream, uint8_t *dev_out,
                               uint64_t *n_classes);

/* `graphdump -f gfa1 | gfa2 | fasta` on the GPU (graphdump.cpp:47-113, 175-582): the compacted
 * de Bruijn graph as text -- segments (non-branching paths between consecutive junction occurrences, ids packed from the
 * junction id, its sign and the first edge character), their occurrences, links and the path of every input sequence.
 *   tpc_graphdump_gfa_device : format 3 (gfa1), 4 (gfa2) or 5 (fasta); the image and the UPPER-CASED characters of all
 *                          input records (back to back, record c = [seq_start[c], seq_start[c + 1]); seq_start and the
 *                          names are host arrays) in device memory.  Produces everything the reference prints after the
 *                          header lines; *dev_text is allocated by the callee (tpc_device_free).
 *   tpc_graphdump_gfa_file   : image file + the FASTA files given with -s -> text file (NULL or "-" = stdout), byte for byte
 *                          what the reference prints (header lines included); prefix != 0 == --prefix.
 * Errors: "The input is corrupted" (graphdump.cpp:461; also for images whose sequences do not match the FASTA files),
 * "A vertex id is too large, cannot generate GFA" (:58). */
int tpc_graphdump_gfa_device(const uint8_t *dev_image, uint64_t image_bytes, uint32_t format, uint32_t k,
                             const uint8_t *dev_seq_chars, const uint64_t *seq_start, const char *const *seq_name,
                             uint64_t n_seq, void *stream, uint8_t **dev_text, uint64_t *text_bytes);
int tpc_graphdump_gfa_file(const char *image_path, const char *format, uint32_t k, const char *const *seq_paths,
                           size_t n_seq_paths, int prefix, const char *out_path);

/* Sessions allocate from the current device's stream-ordered pool, which keeps freed memory for the next run; this
 * hands it back to the driver (e.g. before another process uses the GPU). */
int tpc_release_cached_memory(void);
int tpc_device_alloc(uint64_t bytes, void **out);
void tpc_device_free(void *p);
int tpc_copy_to_host(void *host_dst, const void *dev_src, uint64_t bytes);
int tpc_copy_to_device(void *dev_dst, const void *host_src, uint64_t bytes);

const char *tpc_last_error(void);
uint32_t tpc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TWOPACO_B200_H_ */
