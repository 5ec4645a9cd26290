/* hs_wide.c — the Haskell DO dialects for alphabets of MORE THAN 32 symbols (TEST INFRASTRUCTURE).
 *
 * Same restatement as the `hs_*` section of do_oracle.c (S1 full, S1 Ukkonen, S2 Ukkonen; file references there), with an
 * ambiguity set stored as W = ceil(a / 32) consecutive 32-bit words (bit i of the set = bit i % 32 of word i / 32, the
 * gap is bit a - 1).  measureCharacters compares elements as unsigned naturals (most significant word first), like
 * bv-little's Ord.
 * PINNED end to end by the reference's script tests on the 81-symbol `slashes` data
 * (test/TestSuite/ScriptTests.hs:205-222: discrete 197, 2:1 228, hamming 671, levenshtein 488), tests/test_oracle.py. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "do_oracle.h"

typedef struct { int kind, a, W; const uint32_t *sigma; } wm_t;

static int w_test(const uint32_t *x, int bit) { return (x[bit >> 5] >> (bit & 31)) & 1u; }
static int w_any_and(const uint32_t *x, const uint32_t *y, int W)
{
    for (int w = 0; w < W; w++) if (x[w] & y[w]) return 1;
    return 0;
}
static int w_eq(const uint32_t *x, const uint32_t *y, int W) { return memcmp(x, y, sizeof(uint32_t) * (size_t)W) == 0; }
static int w_cmp(const uint32_t *x, const uint32_t *y, int W)      /* as unsigned naturals */
{
    for (int w = W - 1; w >= 0; w--) if (x[w] != y[w]) return x[w] < y[w] ? -1 : 1;
    return 0;
}

static uint64_t w_overlap(const wm_t *mt, const uint32_t *x, const uint32_t *y, uint32_t *med)
{
    const int W = mt->W, a = mt->a;
    if (mt->kind == 0) {                                     /* discreteMetricPairwiseLogic */
        const int inter = w_any_and(x, y, W);
        for (int w = 0; w < W; w++) med[w] = inter ? (x[w] & y[w]) : (x[w] | y[w]);
        return inter ? 0 : 1;
    }
    if (mt->kind == 2 || mt->kind == 3) {                    /* firstLinearNormPairwiseLogic (MetricRepresentation.hs:166-190) */
        /* toRange (AmbiguityGroup.hs:154-160): (countLeadingZeros, max 0 (bits - countTrailingZeros - 1)), ordered by
         * fromTupleWithPrecision (Range.hs:100-101).  kind 2: bv-little counts "leading" from index 0 (range = lowest..highest
         * set symbol); kind 3: leading = most significant end (range mirrored: a-1-highest .. a-1-lowest). */
        int lo[2], hi[2];
        const uint32_t *e[2] = { x, y };
        for (int k = 0; k < 2; k++) {
            int l = -1, h = -1;
            for (int b = 0; b < a; b++) if (w_test(e[k], b)) { if (l < 0) l = b; h = b; }
            if (l < 0) { l = a; h = -1; }                    /* empty set: clz = ctz = a */
            int first, last;
            if (mt->kind == 2) { first = l; last = (h < 0) ? 0 : h; if (l == a) last = 0; }
            else { first = (h < 0) ? a : a - 1 - h; last = (l == a) ? 0 : a - 1 - l; if (last < 0) last = 0; }
            lo[k] = first < last ? first : last; hi[k] = first < last ? last : first;
        }
        const int isect = lo[1] <= hi[0] && hi[1] >= lo[0];  /* intersects (Range.hs:134-136) */
        int nl, nu;
        if (isect) { nl = lo[0] > lo[1] ? lo[0] : lo[1]; nu = hi[0] < hi[1] ? hi[0] : hi[1]; }          /* intersection */
        else       { nl = hi[0] < hi[1] ? hi[0] : hi[1]; nu = lo[0] > lo[1] ? lo[0] : lo[1]; }          /* smallestClosed */
        memset(med, 0, sizeof(uint32_t) * (size_t)W);        /* fromRange (AmbiguityGroup.hs:162-172), truncated to a bits */
        if (nu == nl) { if (nu < a) med[nu >> 5] = 1u << (nu & 31); }
        else for (int b = nl; b < nu && b < a; b++) med[b >> 5] |= 1u << (b & 31);     /* (2^ub - 1) xor (2^lb - 1) */
        return isect ? 0 : (uint64_t)(nu - nl);
    }
    uint64_t best = UINT64_MAX;
    memset(med, 0, sizeof(uint32_t) * (size_t)W);
    int *xs = malloc(sizeof(int) * 2 * (size_t)a), *ys = xs + a, nx = 0, ny = 0;   /* the set symbols of each input */
    for (int j = 0; j < a; j++) { if (w_test(x, j)) xs[nx++] = j; if (w_test(y, j)) ys[ny++] = j; }
    for (int s = 0; s < a; s++) {                            /* overlap / computeCostMedian: every minimal symbol */
        uint64_t dx = UINT64_MAX, dy = UINT64_MAX;
        const uint32_t *row = mt->sigma + (size_t)s * a;
        for (int k = 0; k < nx; k++) if (row[xs[k]] < dx) dx = row[xs[k]];
        for (int k = 0; k < ny; k++) if (row[ys[k]] < dy) dy = row[ys[k]];
        const uint64_t c = (dx == UINT64_MAX || dy == UINT64_MAX) ? UINT64_MAX : dx + dy;
        if (c < best) { best = c; memset(med, 0, sizeof(uint32_t) * (size_t)W); med[s >> 5] = 1u << (s & 31); }
        else if (c == best && c != UINT64_MAX) med[s >> 5] |= 1u << (s & 31);
    }
    free(xs);
    return best;
}

#define W_INF (UINT64_MAX / 4)
enum { W_DIAG = 0, W_LEFT = 1, W_UP = 2 };

static uint64_t w_fill(const wm_t *mt, int s2, const uint32_t *les, size_t m, const uint32_t *lon, size_t n, long off,
                       const uint32_t *gap, uint8_t *dir, uint64_t *cells)
{
    const int W = mt->W;
    const size_t Wd = n + 1;
    uint64_t *prev = malloc(sizeof(uint64_t) * Wd), *cur = malloc(sizeof(uint64_t) * Wd), cnt = 0;
    uint32_t *md = malloc(sizeof(uint32_t) * (size_t)W);
    for (size_t i = 0; i <= m; i++) {
        for (size_t j = 0; j <= n; j++) {
            const long x = (long)j - (long)i;
            if (x < -off || x > (long)(n - m) + off) { cur[j] = W_INF; continue; }
            cnt++;
            if (i == 0 && j == 0) { cur[0] = 0; dir[0] = W_DIAG; continue; }
            uint64_t best = W_INF; uint8_t bd = W_DIAG; int relabel = 0;
            if (i > 0 && j > 0 && prev[j - 1] < W_INF) {
                const uint64_t c = w_overlap(mt, lon + (j - 1) * W, les + (i - 1) * W, md);
                best = prev[j - 1] + c; bd = W_DIAG; relabel = s2 && w_eq(md, gap, W);
            }
            if (j > 0 && cur[j - 1] < W_INF) {
                const uint64_t c = w_overlap(mt, lon + (j - 1) * W, gap, md);
                if (cur[j - 1] + c < best) { best = cur[j - 1] + c; bd = W_LEFT; relabel = 0; }
            }
            if (i > 0 && prev[j] < W_INF) {
                const uint64_t c = w_overlap(mt, gap, les + (i - 1) * W, md);
                if (prev[j] + c < best) { best = prev[j] + c; bd = W_UP; relabel = 0; }
            }
            cur[j] = best;
            dir[i * Wd + j] = relabel ? W_LEFT : bd;
        }
        uint64_t *t = prev; prev = cur; cur = t;
    }
    const uint64_t res = prev[n];
    free(prev); free(cur); free(md);
    if (cells) *cells += cnt;
    return res;
}

/* as oracle_haskell_do, elements of W = ceil(a/32) words; output arrays need (n1 + n2 + 1) * W words */
int64_t oracle_haskell_do_w(int dialect, int metric_kind, const uint32_t *sigma, int a,
                            const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                            uint32_t *out_med, uint32_t *out_a1, uint32_t *out_a2, size_t *out_len,
                            uint64_t *cells, int *final_offset)
{
    const int W = (a + 31) / 32;
    const wm_t mt = { metric_kind, a, W, sigma };
    uint32_t *gap = calloc((size_t)W, sizeof(uint32_t)), *one = calloc((size_t)W, sizeof(uint32_t)),
             *md = calloc((size_t)W, sizeof(uint32_t));
    gap[(a - 1) >> 5] = 1u << ((a - 1) & 31);
    int swapped = 0;
    if (n1 != n2) swapped = n1 > n2;
    else for (size_t i = 0; i < n1; i++) {
        const int c = w_cmp(in1 + i * W, in2 + i * W, W);
        if (c) { swapped = c > 0; break; }
    }
    const uint32_t *les = swapped ? in2 : in1, *lon = swapped ? in1 : in2;
    const size_t m = swapped ? n2 : n1, n = swapped ? n1 : n2;
    if (cells) *cells = 0;
    if (final_offset) *final_offset = -1;
    uint8_t *dir = calloc((m + 1) * (n + 1), 1);
    uint64_t delta = UINT64_MAX, G = 0;
    for (int s = 0; s < a - 1; s++) {
        memset(one, 0, sizeof(uint32_t) * (size_t)W);
        one[s >> 5] = 1u << (s & 31);
        const uint64_t c1 = w_overlap(&mt, one, gap, md), c2 = w_overlap(&mt, gap, one, md);
        const uint64_t c = c1 < c2 ? c1 : c2;
        if (c < delta) delta = c;
    }
    for (size_t i = 0; i < m; i++) G += w_test(les + i * W, a - 1);
    for (size_t j = 0; j < n; j++) G += w_test(lon + j * W, a - 1);
    const uint64_t qd = n - m + 1;
    int no_gain = m <= 4 || 2 * n >= 3 * m || delta == 0;
    if (dialect == 2) no_gain = no_gain || G >= n;
    uint64_t cost;
    if (dialect == 0 || no_gain) {
        cost = w_fill(&mt, 0, les, m, lon, n, (long)(m + n + 2), gap, dir, cells);
    } else if (dialect == 1) {
        uint64_t off = 2;
        for (;;) {
            cost = w_fill(&mt, 0, les, m, lon, n, (long)(off < m ? off : m), gap, dir, cells);
            const int64_t cv = (int64_t)delta * ((int64_t)qd + (int64_t)off - (int64_t)G);
            const uint64_t thr = cv > 0 ? (uint64_t)cv : 0;
            if (final_offset) *final_offset = (int)off;
            if (thr <= cost) off *= 2; else break;
        }
    } else {
        uint64_t off = 2 + G;
        for (;;) {
            cost = w_fill(&mt, 1, les, m, lon, n, (long)(off < m ? off : m), gap, dir, cells);
            if (final_offset) *final_offset = (int)off;
            if (qd + off > n) break;
            const uint64_t thr = (qd + off <= G) ? 0 : delta * (qd + off - G);
            if (thr <= cost) off *= 2; else break;
        }
    }
    size_t i = m, j = n, k = 0;
    const size_t cap = m + n + 1;
    uint32_t *rm = calloc(cap * W, sizeof(uint32_t)), *rx = calloc(cap * W, sizeof(uint32_t)),
             *ry = calloc(cap * W, sizeof(uint32_t));
    while (i > 0 || j > 0) {
        const uint8_t d = dir[i * (n + 1) + j];
        if (d == W_LEFT)    { w_overlap(&mt, gap, lon + (j - 1) * W, rm + k * W); memcpy(ry + k * W, lon + (j - 1) * W, 4u * W); j--; }
        else if (d == W_UP) { w_overlap(&mt, les + (i - 1) * W, gap, rm + k * W); memcpy(rx + k * W, les + (i - 1) * W, 4u * W); i--; }
        else { w_overlap(&mt, les + (i - 1) * W, lon + (j - 1) * W, rm + k * W);
               memcpy(rx + k * W, les + (i - 1) * W, 4u * W); memcpy(ry + k * W, lon + (j - 1) * W, 4u * W); i--; j--; }
        k++;
    }
    for (size_t t = 0; t < k; t++) {
        memcpy(out_med + t * W, rm + (k - 1 - t) * W, 4u * W);
        memcpy(out_a1 + t * W, (swapped ? ry : rx) + (k - 1 - t) * W, 4u * W);
        memcpy(out_a2 + t * W, (swapped ? rx : ry) + (k - 1 - t) * W, 4u * W);
    }
    *out_len = k;
    free(rm); free(rx); free(ry); free(dir); free(gap); free(one); free(md);
    return (int64_t)cost;
}

/* one overlap lookup on W-word elements */
uint64_t oracle_overlap_w(int metric_kind, const uint32_t *sigma, int a, const uint32_t *x, const uint32_t *y, uint32_t *med)
{
    const wm_t mt = { metric_kind, a, (a + 31) / 32, sigma };
    return w_overlap(&mt, x, y, med);
}
