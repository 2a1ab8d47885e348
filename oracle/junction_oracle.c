/*
 * oracle/junction_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, hash-free CPU restatement of what TwoPaCo's two-pass junction
 * finder computes.  It exists to CHECK the CUDA path (tests/, smoke(), and the
 * cpu_baseline leg of bench.py); nothing in twopaco_b200/ may link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (a) the reference's golden example (example/example.fa, k=11: SURVEY.md
 *       appendix A canonical stream, and the shipped example/example.dbg),
 *   (b) outputs of the unmodified reference binary (oracle/_ref/twopaco built
 *       by oracle/Makefile) on the reference's own --test recipe and on
 *       C2-style sets: tests/golden/golden.json fixtures + live runs when the
 *       binary is present.
 *
 * What is restated (paths relative to /root/reference):
 *   - FASTA framing and alphabet  : src/common/streamfastaparser.cpp:29-133,
 *                                   src/common/dnachar.cpp:9-11,52-85
 *   - non-ACGT -> 'N', sentinels  : src/graphconstructor/vertexenumerator.h:1154,1174,1191
 *   - junction definition         : src/graphconstructor/test.cpp:71-160
 *                                   (== result of the Bloom pass h:995-1105,
 *                                   h:586-704 followed by the exact pass
 *                                   h:708-829, 1228-1256; the Bloom filter
 *                                   only produces a superset, so the exact
 *                                   set is hash-independent)
 *   - id / sign / stub semantics  : src/graphconstructor/bifurcationstorage.h:100-128,
 *                                   vertexenumerator.h:927-958
 *   - output framing              : src/common/junctionapi.h:107-137
 *
 * Ids: the reference's ids depend on /dev/urandom-seeded hashes
 * (mersennetwister.h:242-263); only the partition of occurrences into
 * junctions and the relative strand signs are observable.  This oracle numbers
 * junctions 1..J in order of first appearance with the first occurrence
 * positive, i.e. it emits the canonical relabelling
 * directly; end-of-record stubs get J+42, J+43, ... in stream order
 * (vertexenumerator.h:419,942-948 at -t 1).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ctype.h>

#define ORACLE_MAX_WORDS 20 /* vertexenumerator.h:4 MAX_CAPACITY */

static char g_err[512];
const char *oracle_last_error(void) { return g_err; }

/* ---- alphabet: dnachar.cpp:9-11,18-33,52-85 --------------------------------- */
static int is_definite(int ch) { return ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T'; }
static int is_valid(int ch) { return ch > 0 && strchr("ACGTURYKMSWBDHWNXV", ch) != NULL; }
static int code_of(int ch) /* MakeUpChar */
{
    switch (ch) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return 4;
}

/* ---- FASTA framing: streamfastaparser.cpp:29-93 ----------------------------- */
/* Parses a whole file image.  Returns number of records or -1.  seq[i] is a
 * malloc'd, normalised (upper-case, non-ACGT -> 'N') string of length len[i]. */
typedef struct { char **seq; uint64_t *len; uint64_t n, cap; } rec_list;

static int rec_push(rec_list *r, char *s, uint64_t n)
{
    if (r->n == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 16;
        r->seq = (char **)realloc(r->seq, r->cap * sizeof(char *));
        r->len = (uint64_t *)realloc(r->len, r->cap * sizeof(uint64_t));
        if (!r->seq || !r->len) return -1;
    }
    r->seq[r->n] = s; r->len[r->n] = n; r->n++;
    return 0;
}

int64_t oracle_parse_fasta(const char *path, char ***out_seq, uint64_t **out_len)
{
    FILE *f = fopen(path, "rb");
    if (!f) { snprintf(g_err, sizeof g_err, "Can't open file %s", path); return -1; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    if (sz && fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); snprintf(g_err, sizeof g_err, "read error"); return -1; }
    fclose(f);
    rec_list r = {0};
    long i = 0;
    while (i < sz) {
        /* ReadRecord: first char must be '>' (cpp:31-39), header runs to '\n' */
        if (buf[i] != '>') { snprintf(g_err, sizeof g_err, "The FASTA header should start with a '>', started with '%c'", buf[i]); free(buf); return -1; }
        while (i < sz && buf[i] != '\n') i++;
        if (i < sz) i++;
        /* GetChar loop (cpp:61-93): skip whitespace, stop at '>' anywhere */
        long cap = 1024, n = 0; char *s = (char *)malloc((size_t)cap);
        while (i < sz && buf[i] != '>') {
            int ch = (unsigned char)buf[i++];
            if (isspace(ch)) continue;
            ch = toupper(ch);
            if (!is_valid(ch)) { snprintf(g_err, sizeof g_err, "Found an invalid character '%c'", buf[i - 1]); free(buf); free(s); return -1; }
            if (n == cap) { cap *= 2; s = (char *)realloc(s, (size_t)cap); }
            s[n++] = is_definite(ch) ? (char)ch : 'N'; /* vertexenumerator.h:1174 */
        }
        rec_push(&r, s, (uint64_t)n);
    }
    free(buf);
    *out_seq = r.seq; *out_len = r.len;
    return (int64_t)r.n;
}

void oracle_free_records(char **seq, uint64_t *len, uint64_t n)
{
    for (uint64_t i = 0; i < n; i++) free(seq[i]);
    free(seq); free(len);
}

/* ---- packed k-mers (word 0 most significant, first base most significant) --- */
typedef struct { uint64_t w[ORACLE_MAX_WORDS]; } kmer_t;

static int kmer_cmp(const kmer_t *a, const kmer_t *b, int W)
{
    for (int i = 0; i < W; i++) if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
    return 0;
}

/* append base at the least significant end, drop the most significant base */
static void kmer_push_back(kmer_t *x, int W, uint32_t k, unsigned c)
{
    for (int i = 0; i < W - 1; i++) x->w[i] = (x->w[i] << 2) | (x->w[i + 1] >> 62);
    x->w[W - 1] = (x->w[W - 1] << 2) | c;
    unsigned top = (2 * k) % 64; /* bits used in word 0 */
    if (top) x->w[0] &= (~0ULL) >> (64 - top);
}

/* prepend base at the most significant end, drop the least significant base */
static void kmer_push_front(kmer_t *x, int W, uint32_t k, unsigned c)
{
    for (int i = W - 1; i > 0; i--) x->w[i] = (x->w[i] >> 2) | (x->w[i - 1] << 62);
    x->w[0] >>= 2;
    unsigned top = (2 * k) % 64;
    unsigned sh = top ? top - 2 : 62;
    x->w[0] |= (uint64_t)c << sh;
}

static uint64_t kmer_hash(const kmer_t *x, int W)
{
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (int i = 0; i < W; i++) {
        h ^= x->w[i]; h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    }
    return h;
}

/* ---- vertex table: canonical k-mer -> neighbour sets (test.cpp:110-134) ------ */
typedef struct {
    uint64_t count;      /* occurrences, both strands (candidateoccurence.h Inc) */
    int64_t id;          /* 0 = not yet numbered */
    uint8_t in_mask, out_mask; /* definite neighbours seen, canonical orientation */
    uint8_t in_n, out_n; /* 'N'/virtual neighbours seen, saturating at 2 (each is unique: test.cpp:80-93) */
    uint8_t first_fwd;   /* strand of the first occurrence: 1 = as written, 0 = reverse complement */
    uint8_t used;
} vinfo;

typedef struct { vinfo *v; uint64_t *keys; uint64_t cap, n; int W; } vtable;

static int vt_init(vtable *t, int W, uint64_t cap)
{
    t->W = W; t->cap = cap; t->n = 0;
    t->v = (vinfo *)calloc(cap, sizeof(vinfo));
    t->keys = (uint64_t *)malloc(cap * (size_t)W * 8);
    return (t->v && t->keys) ? 0 : -1;
}

static vinfo *vt_find(vtable *t, const kmer_t *key, int insert);

static int vt_grow(vtable *t)
{
    vtable nt;
    if (vt_init(&nt, t->W, t->cap * 2)) return -1;
    for (uint64_t i = 0; i < t->cap; i++) if (t->v[i].used) {
        kmer_t k; memcpy(k.w, t->keys + i * t->W, (size_t)t->W * 8);
        vinfo *d = vt_find(&nt, &k, 1); *d = t->v[i];
    }
    free(t->v); free(t->keys); *t = nt;
    return 0;
}

static vinfo *vt_find(vtable *t, const kmer_t *key, int insert)
{
    if (insert && (t->n + 1) * 10 > t->cap * 6) { if (vt_grow(t)) return NULL; }
    uint64_t i = kmer_hash(key, t->W) & (t->cap - 1);
    for (;;) {
        if (!t->v[i].used) {
            if (!insert) return NULL;
            memcpy(t->keys + i * t->W, key->w, (size_t)t->W * 8);
            memset(&t->v[i], 0, sizeof(vinfo)); t->v[i].used = 1; t->n++;
            return &t->v[i];
        }
        if (memcmp(t->keys + i * t->W, key->w, (size_t)t->W * 8) == 0) return &t->v[i];
        i = (i + 1) & (t->cap - 1);
    }
}

static int is_junction(const vinfo *e, uint64_t abundance)
{
    int indeg = __builtin_popcount(e->in_mask) + e->in_n;
    int outdeg = __builtin_popcount(e->out_mask) + e->out_n;
    /* test.cpp:136-148 (|in|>1 or |out|>1) and vertexenumerator.h:1239 (Count <= abundance) */
    return (indeg > 1 || outdeg > 1) && e->count <= abundance;
}

/* ---- output buffer: junctionapi.h:107-137 ---------------------------------- */
typedef struct { uint8_t *p; uint64_t n, cap; uint32_t now_chr; } obuf;

static int ob_put(obuf *o, uint32_t pos, int64_t id)
{
    if (o->n + 12 > o->cap) { o->cap = o->cap ? o->cap * 2 : 4096; o->p = (uint8_t *)realloc(o->p, o->cap); if (!o->p) return -1; }
    memcpy(o->p + o->n, &pos, 4); memcpy(o->p + o->n + 4, &id, 8); o->n += 12;
    return 0;
}

static int ob_write_junction(obuf *o, uint32_t chr, uint32_t pos, int64_t id)
{
    for (; chr > o->now_chr; ++o->now_chr) if (ob_put(o, 0xFFFFFFFFu, INT64_MAX)) return -1; /* :120-123 */
    return ob_put(o, pos, id);
}

/*
 * Walk every definite k-mer occurrence of one record.  cb(ctx, pos, fwd, rc,
 * prev, next): prev/next are codes 0..3 or 4 for 'N' / virtual sentinel.
 */
typedef void (*occ_cb)(void *ctx, uint32_t seq, uint64_t pos, const kmer_t *fwd, const kmer_t *rc, int prev, int next);

static void walk_record(const char *s, uint64_t n, uint32_t seq, uint32_t k, int W, occ_cb cb, void *ctx)
{
    if (n < k) return; /* vertexenumerator.h:1177: no task is ever produced */
    kmer_t f, r; memset(&f, 0, sizeof f); memset(&r, 0, sizeof r);
    uint64_t run = 0; /* consecutive definite bases ending here (definiteCount, h:632,666) */
    for (uint64_t i = 0; i < n; i++) {
        int c = code_of(s[i]);
        if (c < 4) { kmer_push_back(&f, W, k, (unsigned)c); kmer_push_front(&r, W, k, (unsigned)(3 - c)); run++; }
        else { run = 0; memset(&f, 0, sizeof f); memset(&r, 0, sizeof r); }
        if (run >= k) {
            uint64_t pos = i + 1 - k;
            int prev = pos == 0 ? 4 : code_of(s[pos - 1]);        /* sentinel 'N' h:1154 */
            int next = pos + k >= n ? 4 : code_of(s[pos + k]);    /* sentinel 'N' h:1191 */
            cb(ctx, seq, pos, &f, &r, prev, next);
        }
    }
}

typedef struct { vtable *t; int W; int fail; } count_ctx;

static void count_cb(void *vctx, uint32_t seq, uint64_t pos, const kmer_t *fwd, const kmer_t *rc, int prev, int next)
{
    (void)seq; (void)pos;
    count_ctx *c = (count_ctx *)vctx;
    int fwd_is_canon = kmer_cmp(fwd, rc, c->W) <= 0;
    /* neighbours in the canonical orientation (candidateoccurence.h:34-47) */
    int in = fwd_is_canon ? prev : (next == 4 ? 4 : 3 - next);
    int out = fwd_is_canon ? next : (prev == 4 ? 4 : 3 - prev);
    vinfo *e = vt_find(c->t, fwd_is_canon ? fwd : rc, 1);
    if (!e) { c->fail = 1; return; }
    if (e->count == 0) e->first_fwd = (uint8_t)fwd_is_canon;
    e->count++;
    if (in < 4) e->in_mask |= (uint8_t)(1 << in); else if (e->in_n < 2) e->in_n++;
    if (out < 4) e->out_mask |= (uint8_t)(1 << out); else if (e->out_n < 2) e->out_n++;
}

typedef struct {
    vtable *t; int W; uint32_t k; uint64_t abundance; obuf *o; const uint64_t *len;
    int64_t next_id; int64_t next_stub; uint64_t marks; uint64_t n_junction_marks; int fail;
    uint8_t *markvec; /* optional: per-record boolean marks for the current record */
    int pass; /* 0 = number junctions only, 1 = emit */
} emit_ctx;

static void emit_cb(void *vctx, uint32_t seq, uint64_t pos, const kmer_t *fwd, const kmer_t *rc, int prev, int next)
{
    (void)prev; (void)next;
    emit_ctx *c = (emit_ctx *)vctx;
    int fwd_is_canon = kmer_cmp(fwd, rc, c->W) <= 0;
    vinfo *e = vt_find(c->t, fwd_is_canon ? fwd : rc, 0);
    int64_t id = INT64_MAX; /* INVALID_VERTEX, common.cpp:5 */
    if (e && is_junction(e, c->abundance)) {
        if (e->id == 0) e->id = c->next_id++;
        /* +id when this occurrence is on the strand of the stored key
         * (bifurcationstorage.h:100-128); our stored strand = first occurrence */
        id = ((uint8_t)fwd_is_canon == e->first_fwd) ? e->id : -e->id;
    }
    if (c->pass == 0) return;
    uint64_t n = c->len[seq];
    int is_end = (pos == 0) || (pos == n - c->k);                 /* h:942 */
    if (id != INT64_MAX) {
        c->marks++; c->n_junction_marks++;
        if (ob_write_junction(c->o, seq, (uint32_t)pos, id)) c->fail = 1; /* h:938 (u32 truncation) */
    } else if (is_end) {
        c->marks++;
        if (ob_write_junction(c->o, seq, (uint32_t)pos, c->next_stub++)) c->fail = 1; /* h:942-948 */
    }
}

/* The first/last k-mer of a record is emitted even when it contains an 'N'
 * (the reference tests definiteCount only for the junction lookup, h:932, not
 * for the stub branch, h:942); walk_record only visits definite k-mers, so
 * handle the non-definite ends here. */
static int window_definite(const char *s, uint64_t pos, uint32_t k)
{
    for (uint32_t i = 0; i < k; i++) if (!is_definite(s[pos + i])) return 0;
    return 1;
}

/*
 * seqs: normalised records (chars in ACGTN; anything else is treated as N).
 * On success returns 0 and malloc's *out (de_bruijn.bin image).
 */
int oracle_find_junctions(const char *const *seqs, const uint64_t *lens, uint64_t nseq, uint32_t k,
                          uint64_t abundance, uint8_t **out, uint64_t *out_bytes,
                          uint64_t *n_junctions, uint64_t *n_marks)
{
    if (k == 0 || (k % 2) == 0) { snprintf(g_err, sizeof g_err, "value of K must be odd"); return 1; }
    int W = (int)((2ULL * k + 63) / 64);
    if (W > ORACLE_MAX_WORDS) { snprintf(g_err, sizeof g_err, "The value of K is too big"); return 1; }
    vtable t;
    if (vt_init(&t, W, 1 << 16)) { snprintf(g_err, sizeof g_err, "out of memory"); return 1; }
    count_ctx cc = { &t, W, 0 };
    for (uint64_t s = 0; s < nseq; s++) walk_record(seqs[s], lens[s], (uint32_t)s, k, W, count_cb, &cc);
    if (cc.fail) { snprintf(g_err, sizeof g_err, "out of memory"); return 1; }

    uint64_t J = 0;
    for (uint64_t i = 0; i < t.cap; i++) if (t.v[i].used && is_junction(&t.v[i], abundance)) J++;

    obuf o = {0};
    emit_ctx ec; memset(&ec, 0, sizeof ec);
    ec.t = &t; ec.W = W; ec.k = k; ec.abundance = abundance; ec.o = &o; ec.len = lens;
    ec.next_id = 1; ec.next_stub = (int64_t)J + 42; ec.pass = 1;
    for (uint64_t s = 0; s < nseq; s++) {
        uint64_t n = lens[s];
        if (n < k) continue;
        /* A non-definite first k-mer is emitted before any definite occurrence
         * of the record; a non-definite last one after all of them. */
        if (!window_definite(seqs[s], 0, k)) {
            ec.marks++;
            if (ob_write_junction(&o, (uint32_t)s, 0, ec.next_stub++)) ec.fail = 1;
        }
        walk_record(seqs[s], n, (uint32_t)s, k, W, emit_cb, &ec);
        if (n > k && !window_definite(seqs[s], n - k, k)) {
            ec.marks++;
            if (ob_write_junction(&o, (uint32_t)s, (uint32_t)(n - k), ec.next_stub++)) ec.fail = 1;
        }
    }
    free(t.v); free(t.keys);
    if (ec.fail) { free(o.p); snprintf(g_err, sizeof g_err, "out of memory"); return 1; }
    *out = o.p; *out_bytes = o.n;
    if (n_junctions) *n_junctions = J;
    if (n_marks) *n_marks = ec.marks;
    return 0;
}

void oracle_free(void *p) { free(p); }
