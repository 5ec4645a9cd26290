/* cpu_bench.c — times the CPU implementations of the path on host cores (TEST / BENCH INFRASTRUCTURE).
 *
 * kind "reference": dlopen()s oracle/_ref/libpcgdo_ref.so — the UNMODIFIED reference aligner — and calls
 *   its align2d exactly as the Haskell binding does (DO/Pairwise/FFI.hsc:272-290, 474-515): four malloc'ed
 *   alignIO_t buffers of capacity len1+len2+2 per call, payload right-aligned, freed after the call.
 * kind "port": the restatement in do_oracle.c.
 * Pairs are split over `threads` pthreads (the entry points are re-entrant: call-local allocations and
 * a read-only cost matrix).  Only bench.py's cpu_baseline / --impl reference legs and tests call this.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "do_oracle.h"

/* zlib-compatible CRC-32 (reflected 0xEDB88320), so that tests can compare whole output strings of many pairs
 * through zlib.crc32 on their side */
static uint32_t crc_table[256];
static void crc_init(void)
{
    if (crc_table[1]) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[i] = c;
    }
}
static uint32_t crc_u8_of_u32(const unsigned int *v, size_t n)      /* elements as ONE byte each */
{
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; i++) c = crc_table[(c ^ (uint8_t)v[i]) & 0xff] ^ (c >> 8);
    return c ^ 0xffffffffu;
}
static uint32_t crc_bytes(const void *p, size_t n)
{
    const uint8_t *b = p;
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; i++) c = crc_table[(c ^ b[i]) & 0xff] ^ (c >> 8);
    return c ^ 0xffffffffu;
}

typedef struct { unsigned int *character; size_t length; size_t capacity; } aio_t;
typedef int (*align2d_fn)(aio_t *, aio_t *, aio_t *, aio_t *, void *, int, int, int);
typedef int (*align2daff_fn)(aio_t *, aio_t *, aio_t *, aio_t *, void *, int);
typedef void (*setup_fn)(void *, unsigned int *, size_t, unsigned int);

typedef struct {
    int tid, threads;
    size_t n;
    const uint8_t *elems;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    int32_t *cost;
    uint64_t *cells;
    align2d_fn ref_align;
    void *ref_cm;
    const oracle_tcm_t *port;
    align2daff_fn ref_affine;      /* non-NULL: affine dialect through the reference */
    int affine_variant;            /* port, affine: 0 as coded, 1 patched; -1 = linear */
    uint32_t *out_len;             /* optional [n]: alignment length */
    uint32_t *crc;                 /* optional [n][3]: CRC-32 of gappedOutput, inputChar1, inputChar2 after the call */
} job_t;

static void fill_aio(aio_t *io, const uint8_t *src, size_t n, size_t cap)
{
    io->character = calloc(cap, sizeof(unsigned int));
    io->length = n;
    io->capacity = cap;
    for (size_t i = 0; i < n; i++) io->character[cap - n + i] = src[i];
}

static void *worker(void *arg)
{
    job_t *j = arg;
    for (size_t k = j->tid; k < j->n; k += j->threads) {
        const uint8_t *s1 = j->elems + j->off1[k], *s2 = j->elems + j->off2[k];
        const size_t n1 = j->len1[k], n2 = j->len2[k];
        if (j->ref_align) {
            const size_t cap = n1 + n2 + 2;
            aio_t a, b, g, u;
            fill_aio(&a, s1, n1, cap);
            fill_aio(&b, s2, n2, cap);
            fill_aio(&g, NULL, 0, cap);
            fill_aio(&u, NULL, 0, cap);
            j->cost[k] = j->ref_affine ? j->ref_affine(&a, &b, &g, &u, j->ref_cm, 1)
                                       : j->ref_align(&a, &b, &g, &u, j->ref_cm, 0, 1, 0);
            if (j->out_len) j->out_len[k] = (uint32_t)g.length;
            if (j->crc) {
                j->crc[3 * k + 0] = crc_u8_of_u32(g.character + (g.capacity - g.length), g.length);
                j->crc[3 * k + 1] = crc_u8_of_u32(a.character + (a.capacity - a.length), a.length);
                j->crc[3 * k + 2] = crc_u8_of_u32(b.character + (b.capacity - b.length), b.length);
            }
            free(a.character); free(b.character); free(g.character); free(u.character);
            if (j->cells && !j->ref_affine) {
                size_t lg = (n1 > n2 ? n1 : n2) + 1, sh = (n1 > n2 ? n2 : n1) + 1;
                j->cells[k] = oracle_band_rows(lg, sh, NULL, NULL, NULL, NULL);
            }
        } else {
            uint32_t *buf = malloc(sizeof(uint32_t) * (5 * (n1 + n2) + 8));
            uint32_t *x1 = buf, *x2 = buf + n1, *og = x2 + n2, *o1 = og + n1 + n2 + 2, *o2 = o1 + n1 + n2 + 2;
            for (size_t i = 0; i < n1; i++) x1[i] = s1[i];
            for (size_t i = 0; i < n2; i++) x2[i] = s2[i];
            size_t nal = 0;
            uint64_t c = 0;
            if (j->affine_variant >= 0)
                j->cost[k] = (int32_t)oracle_align2d_affine(j->port, j->affine_variant, x1, n1, x2, n2, og, o1, o2, &nal, &c);
            else
                j->cost[k] = (int32_t)oracle_align2d(j->port, x1, n1, x2, n2, og, o1, o2, &nal, &c);
            if (j->cells) j->cells[k] = c;
            if (j->out_len) j->out_len[k] = (uint32_t)nal;
            if (j->crc) {
                j->crc[3 * k + 0] = crc_u8_of_u32(og, nal);
                j->crc[3 * k + 1] = crc_u8_of_u32(o1, nal);
                j->crc[3 * k + 2] = crc_u8_of_u32(o2, nal);
            }
            free(buf);
        }
    }
    return NULL;
}

/* Returns 0 on success; *seconds = wall time of the alignment loop only (input already in memory). */
int cpu_bench_align2d_ex(const char *ref_so, const uint32_t *sigma, int a, uint32_t gap_open, int affine_variant,
                         const uint8_t *elems, const uint64_t *off1, const uint32_t *len1,
                         const uint64_t *off2, const uint32_t *len2, size_t n, int threads,
                         int32_t *cost, uint64_t *cells, double *seconds);

int cpu_bench_align2d(const char *ref_so, const uint32_t *sigma, int a,
                      const uint8_t *elems, const uint64_t *off1, const uint32_t *len1,
                      const uint64_t *off2, const uint32_t *len2, size_t n, int threads,
                      int32_t *cost, uint64_t *cells, double *seconds)
{
    return cpu_bench_align2d_ex(ref_so, sigma, a, 0, -1, elems, off1, len1, off2, len2, n, threads, cost, cells, seconds);
}

int cpu_bench_align2d_full(const char *ref_so, const uint32_t *sigma, int a, uint32_t gap_open, int affine_variant,
                           const uint8_t *elems, const uint64_t *off1, const uint32_t *len1,
                           const uint64_t *off2, const uint32_t *len2, size_t n, int threads,
                           int32_t *cost, uint64_t *cells, uint32_t *out_len, uint32_t *crc, double *seconds);

int cpu_bench_align2d_ex(const char *ref_so, const uint32_t *sigma, int a, uint32_t gap_open, int affine_variant,
                         const uint8_t *elems, const uint64_t *off1, const uint32_t *len1,
                         const uint64_t *off2, const uint32_t *len2, size_t n, int threads,
                         int32_t *cost, uint64_t *cells, double *seconds)
{
    return cpu_bench_align2d_full(ref_so, sigma, a, gap_open, affine_variant, elems, off1, len1, off2, len2, n, threads,
                                  cost, cells, NULL, NULL, seconds);
}

/* gap_open > 0: the affine dialect (reference align2dAffine, or the port with affine_variant 0 / 1).
 * out_len / crc (optional): per pair the alignment length and the CRC-32 of the three output strings — whole-string
 * parity of large batches against the compiled reference without shipping the strings. */
int cpu_bench_align2d_full(const char *ref_so, const uint32_t *sigma, int a, uint32_t gap_open, int affine_variant,
                           const uint8_t *elems, const uint64_t *off1, const uint32_t *len1,
                           const uint64_t *off2, const uint32_t *len2, size_t n, int threads,
                           int32_t *cost, uint64_t *cells, uint32_t *out_len, uint32_t *crc, double *seconds)
{
    crc_init();
    align2d_fn ref_align = NULL;
    align2daff_fn ref_affine = NULL;
    void *cm = NULL;
    oracle_tcm_t *port = NULL;
    if (ref_so) {
        void *h = dlopen(ref_so, RTLD_NOW | RTLD_LOCAL);
        if (!h) return 1;
        ref_align = (align2d_fn)dlsym(h, "align2d");
        if (gap_open) ref_affine = (align2daff_fn)dlsym(h, "align2dAffine");
        setup_fn setup = (setup_fn)dlsym(h, "setUp2dCostMtx");
        if (!ref_align || !setup || (gap_open && !ref_affine)) return 2;
        cm = calloc(1, 256);
        uint32_t *sg = malloc(sizeof(uint32_t) * a * a);
        memcpy(sg, sigma, sizeof(uint32_t) * a * a);
        setup(cm, sg, (size_t)a, gap_open);
        free(sg);
    } else {
        port = oracle_tcm_new(sigma, a, gap_open);
    }
    if (!gap_open) affine_variant = -1;
    if (threads < 1) threads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * threads);
    job_t *jobs = malloc(sizeof(job_t) * threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (job_t){ t, threads, n, elems, off1, off2, len1, len2, cost, cells, ref_align, cm, port, ref_affine, affine_variant,
                           out_len, crc };
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(th); free(jobs);
    if (port) oracle_tcm_free(port);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * CPU post-order loop (BASELINE.md config 5 baseline): for every (tree, character) column, children before
 * parents, node string = the gapped median of align2d(left, right) with its gap elements removed (what the
 * Haskell wrapper's deleteGaps feeds to the next alignment), total = local + children.
 * Columns are split over `threads` pthreads.  ref_so != NULL -> the unmodified reference align2d.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int tid, threads;
    uint32_t n_trees, n_taxa, n_chars, col0, n_cols;   /* columns col0 .. col0+n_cols-1 of the tree x char grid */
    const int32_t *children;
    const uint8_t *leaf_elems;
    const uint64_t *leaf_off;
    const uint32_t *leaf_len;
    int64_t *col_cost;         /* [n_cols] root total */
    uint64_t *col_cells;       /* [n_cols] */
    align2d_fn ref_align;
    void *ref_cm;
    const oracle_tcm_t *port;
    uint32_t gap;
    int64_t *node_total;       /* optional [n_cols][n_taxa-1]: total cost of every internal node */
    int32_t *node_local;       /* optional, same shape: local cost */
    uint32_t *node_crc;        /* optional, same shape: CRC-32 of the node's ungapped median string */
} po_job_t;

static void *po_worker(void *arg)
{
    po_job_t *j = arg;
    const uint32_t ni = j->n_taxa - 1, nn = 2 * j->n_taxa - 1;
    uint8_t **str = calloc(nn, sizeof(uint8_t *));
    uint32_t *len = calloc(nn, sizeof(uint32_t));
    int64_t *tot = calloc(nn, sizeof(int64_t));
    int *done = calloc(nn, sizeof(int));
    for (uint32_t col = j->tid; col < j->n_cols; col += j->threads) {
        const uint32_t g = j->col0 + col, t = g / j->n_chars, c = g % j->n_chars;
        const int32_t *ch = j->children + (size_t)t * ni * 2;
        uint64_t cells = 0;
        for (uint32_t v = 0; v < nn; v++) { done[v] = v < j->n_taxa; tot[v] = 0; }
        for (uint32_t v = 0; v < j->n_taxa; v++) {
            str[v] = (uint8_t *)j->leaf_elems + j->leaf_off[(size_t)v * j->n_chars + c];
            len[v] = j->leaf_len[(size_t)v * j->n_chars + c];
        }
        uint32_t remaining = ni, root = 0;
        while (remaining) {
            for (uint32_t x = 0; x < ni; x++) {
                const uint32_t node = j->n_taxa + x;
                if (done[node] || !done[ch[2 * x]] || !done[ch[2 * x + 1]]) continue;
                const uint32_t l = ch[2 * x], r = ch[2 * x + 1];
                const size_t n1 = len[l], n2 = len[r], cap = n1 + n2 + 2;
                uint8_t *out = malloc(cap);
                uint32_t no = 0;
                int64_t cost;
                if (j->ref_align) {
                    aio_t a, b, gg, u;
                    fill_aio(&a, str[l], n1, cap);
                    fill_aio(&b, str[r], n2, cap);
                    fill_aio(&gg, NULL, 0, cap);
                    fill_aio(&u, NULL, 0, cap);
                    cost = j->ref_align(&a, &b, &gg, &u, j->ref_cm, 0, 1, 0);
                    for (size_t k = 0; k < gg.length; k++) {
                        unsigned int m = gg.character[gg.capacity - gg.length + k];
                        if (m != j->gap) out[no++] = (uint8_t)m;
                    }
                    free(a.character); free(b.character); free(gg.character); free(u.character);
                } else {
                    uint32_t *buf = malloc(sizeof(uint32_t) * (5 * cap + 8));
                    uint32_t *x1 = buf, *x2 = buf + n1, *og = x2 + n2, *o1 = og + cap, *o2 = o1 + cap;
                    for (size_t i = 0; i < n1; i++) x1[i] = str[l][i];
                    for (size_t i = 0; i < n2; i++) x2[i] = str[r][i];
                    size_t nal = 0;
                    uint64_t cc = 0;
                    cost = oracle_align2d(j->port, x1, n1, x2, n2, og, o1, o2, &nal, &cc);
                    for (size_t k = 0; k < nal; k++)
                        if (og[k] != j->gap) out[no++] = (uint8_t)og[k];
                    free(buf);
                }
                {
                    size_t lg = (n1 > n2 ? n1 : n2) + 1, sh = (n1 > n2 ? n2 : n1) + 1;
                    cells += oracle_band_rows(lg, sh, NULL, NULL, NULL, NULL);
                }
                str[node] = out;
                len[node] = no;
                tot[node] = cost + tot[l] + tot[r];
                if (j->node_total) j->node_total[(size_t)col * ni + x] = tot[node];
                if (j->node_local) j->node_local[(size_t)col * ni + x] = (int32_t)cost;
                if (j->node_crc) j->node_crc[(size_t)col * ni + x] = crc_bytes(out, no);
                done[node] = 1;
                root = node;
                remaining--;
            }
        }
        j->col_cost[col] = tot[root];        /* the last node completed is the root */
        if (j->col_cells) j->col_cells[col] = cells;
        for (uint32_t v = j->n_taxa; v < nn; v++) { free(str[v]); str[v] = NULL; }
    }
    free(str); free(len); free(tot); free(done);
    return NULL;
}

int cpu_bench_postorder_full(const char *ref_so, const uint32_t *sigma, int a, uint32_t n_trees, uint32_t n_taxa,
                             uint32_t n_chars, const int32_t *children, const uint8_t *leaf_elems,
                             const uint64_t *leaf_off, const uint32_t *leaf_len, uint32_t col0, uint32_t n_cols,
                             int threads, int64_t *col_cost, uint64_t *col_cells, int64_t *node_total, int32_t *node_local,
                             uint32_t *node_crc, double *seconds);

int cpu_bench_postorder(const char *ref_so, const uint32_t *sigma, int a, uint32_t n_trees, uint32_t n_taxa,
                        uint32_t n_chars, const int32_t *children, const uint8_t *leaf_elems,
                        const uint64_t *leaf_off, const uint32_t *leaf_len, uint32_t col0, uint32_t n_cols,
                        int threads, int64_t *col_cost, uint64_t *col_cells, double *seconds)
{
    return cpu_bench_postorder_full(ref_so, sigma, a, n_trees, n_taxa, n_chars, children, leaf_elems, leaf_off, leaf_len,
                                    col0, n_cols, threads, col_cost, col_cells, NULL, NULL, NULL, seconds);
}

/* node_total / node_local / node_crc (optional, [n_cols][n_taxa - 1]): every internal node's decoration, for parity
 * of a whole forest column set against the compiled reference. */
int cpu_bench_postorder_full(const char *ref_so, const uint32_t *sigma, int a, uint32_t n_trees, uint32_t n_taxa,
                             uint32_t n_chars, const int32_t *children, const uint8_t *leaf_elems,
                             const uint64_t *leaf_off, const uint32_t *leaf_len, uint32_t col0, uint32_t n_cols,
                             int threads, int64_t *col_cost, uint64_t *col_cells, int64_t *node_total, int32_t *node_local,
                             uint32_t *node_crc, double *seconds)
{
    crc_init();
    align2d_fn ref_align = NULL;
    void *cm = NULL;
    oracle_tcm_t *port = NULL;
    if (ref_so) {
        void *h = dlopen(ref_so, RTLD_NOW | RTLD_LOCAL);
        if (!h) return 1;
        ref_align = (align2d_fn)dlsym(h, "align2d");
        setup_fn setup = (setup_fn)dlsym(h, "setUp2dCostMtx");
        if (!ref_align || !setup) return 2;
        cm = calloc(1, 256);
        uint32_t *sg = malloc(sizeof(uint32_t) * a * a);
        memcpy(sg, sigma, sizeof(uint32_t) * a * a);
        setup(cm, sg, (size_t)a, 0);
        free(sg);
    } else {
        port = oracle_tcm_new(sigma, a, 0);
    }
    if (threads < 1) threads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * threads);
    po_job_t *jobs = malloc(sizeof(po_job_t) * threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (po_job_t){ t, threads, n_trees, n_taxa, n_chars, col0, n_cols, children, leaf_elems, leaf_off,
                              leaf_len, col_cost, col_cells, ref_align, cm, port, 1u << (a - 1), node_total, node_local,
                              node_crc };
        pthread_create(&th[t], NULL, po_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(th); free(jobs);
    if (port) oracle_tcm_free(port);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Large-alphabet dialects (BASELINE config 4 baseline).  The reference runs these in Haskell
 * (unboxedUkkonenFullSpaceDO); no GHC here, so the timed code is the restatement in do_oracle.c ("port").
 * memo_so != NULL: every per-cell overlap goes through the UNMODIFIED compiled reference tcm-memo
 * (getCostAndMedian2D with freshly malloc'ed dcElement_t buffers per call, as
 * lib/tcm-memo/src/Data/TCM/Memoized/FFI.hsc:97-112 does) — one memo matrix per thread, because the reference's
 * unordered_map is not thread-safe.  memo_so == NULL: metric_kind 0 (discrete closed form) or 1 (direct Sankoff).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { size_t alphSize; uint64_t *element; } dce_t;
typedef void *(*minit_fn)(size_t, unsigned int *);
typedef unsigned int (*mget_fn)(dce_t *, dce_t *, dce_t *, void *);

typedef struct { mget_fn get; void *matrix; size_t a; uint64_t lookups; } memo_ctx_t;

static uint64_t memo_lookup(void *vctx, uint32_t x, uint32_t y, uint32_t *med)
{
    memo_ctx_t *c = vctx;
    dce_t e1 = { c->a, calloc(1, 8) }, e2 = { c->a, calloc(1, 8) }, r = { c->a, calloc(1, 8) };
    e1.element[0] = x; e2.element[0] = y;
    const unsigned int cost = c->get(&e1, &e2, &r, c->matrix);
    *med = (uint32_t)r.element[0];
    free(e1.element); free(e2.element); free(r.element);
    c->lookups++;
    return cost;
}

typedef struct {
    int tid, threads, dialect, metric_kind, a;
    const uint32_t *sigma;
    size_t n;
    const uint32_t *elems;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    int32_t *cost;
    uint64_t *cells;
    memo_ctx_t *memo;      /* per thread, or NULL */
    uint32_t *out_len;     /* optional [n] */
    uint32_t *crc;         /* optional [n][3]: CRC-32 over the little-endian words of median, aligned first, aligned second */
    int32_t *final_offset; /* optional [n] */
} hs_job_t;

static void *hs_worker(void *arg)
{
    hs_job_t *j = arg;
    for (size_t k = j->tid; k < j->n; k += j->threads) {
        const uint32_t *s1 = j->elems + j->off1[k], *s2 = j->elems + j->off2[k];
        const size_t n1 = j->len1[k], n2 = j->len2[k], cap = n1 + n2 + 2;
        uint32_t *buf = malloc(sizeof(uint32_t) * 3 * cap);
        size_t nal = 0;
        uint64_t c = 0;
        int off = 0;
        if (j->memo)
            j->cost[k] = (int32_t)oracle_haskell_do_lookup(j->dialect, memo_lookup, j->memo, j->a, s1, n1, s2, n2,
                                                           buf, buf + cap, buf + 2 * cap, &nal, &c, &off);
        else
            j->cost[k] = (int32_t)oracle_haskell_do(j->dialect, j->metric_kind, j->sigma, j->a, s1, n1, s2, n2,
                                                    buf, buf + cap, buf + 2 * cap, &nal, &c, &off);
        if (j->cells) j->cells[k] = c;
        if (j->out_len) j->out_len[k] = (uint32_t)nal;
        if (j->final_offset) j->final_offset[k] = off;
        if (j->crc)
            for (int x = 0; x < 3; x++) j->crc[3 * k + x] = crc_bytes(buf + x * cap, 4 * nal);
        free(buf);
    }
    return NULL;
}

int cpu_bench_haskell_full(const char *memo_so, int dialect, int metric_kind, const uint32_t *sigma, int a,
                           const uint32_t *elems, const uint64_t *off1, const uint32_t *len1,
                           const uint64_t *off2, const uint32_t *len2, size_t n, int threads,
                           int32_t *cost, uint64_t *cells, uint32_t *out_len, uint32_t *crc, int32_t *final_offset,
                           double *seconds, uint64_t *lookups);

int cpu_bench_haskell(const char *memo_so, int dialect, int metric_kind, const uint32_t *sigma, int a,
                      const uint32_t *elems, const uint64_t *off1, const uint32_t *len1,
                      const uint64_t *off2, const uint32_t *len2, size_t n, int threads,
                      int32_t *cost, uint64_t *cells, double *seconds, uint64_t *lookups)
{
    return cpu_bench_haskell_full(memo_so, dialect, metric_kind, sigma, a, elems, off1, len1, off2, len2, n, threads, cost,
                                  cells, NULL, NULL, NULL, seconds, lookups);
}

int cpu_bench_haskell_full(const char *memo_so, int dialect, int metric_kind, const uint32_t *sigma, int a,
                           const uint32_t *elems, const uint64_t *off1, const uint32_t *len1,
                           const uint64_t *off2, const uint32_t *len2, size_t n, int threads,
                           int32_t *cost, uint64_t *cells, uint32_t *out_len, uint32_t *crc, int32_t *final_offset,
                           double *seconds, uint64_t *lookups)
{
    crc_init();
    if (threads < 1) threads = 1;
    memo_ctx_t *memos = NULL;
    if (memo_so) {
        void *h = dlopen(memo_so, RTLD_NOW | RTLD_LOCAL);
        if (!h) return 1;
        minit_fn init = (minit_fn)dlsym(h, "matrixInit");
        mget_fn get = (mget_fn)dlsym(h, "getCostAndMedian2D");
        if (!init || !get) return 2;
        memos = calloc(threads, sizeof(memo_ctx_t));
        for (int t = 0; t < threads; t++) {
            unsigned int *sg = malloc(sizeof(unsigned int) * a * a);
            memcpy(sg, sigma, sizeof(unsigned int) * a * a);
            memos[t] = (memo_ctx_t){ get, init((size_t)a, sg), (size_t)a, 0 };
            free(sg);
        }
    }
    pthread_t *th = malloc(sizeof(pthread_t) * threads);
    hs_job_t *jobs = malloc(sizeof(hs_job_t) * threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (hs_job_t){ t, threads, dialect, metric_kind, a, sigma, n, elems, off1, off2, len1, len2, cost, cells,
                              memos ? &memos[t] : NULL, out_len, crc, final_offset };
        pthread_create(&th[t], NULL, hs_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (lookups) {
        *lookups = 0;
        for (int t = 0; memos && t < threads; t++) *lookups += memos[t].lookups;
    }
    free(th); free(jobs); free(memos);     /* the memo matrices are leaked, like the reference does */
    return 0;
}
