/* do_oracle.c — plain-C CPU restatement of PCG's pairwise DO path (see do_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path and the "port" CPU baseline.
 * Written from the reference's behaviour, not copied: one straightforward full-size
 * direction matrix, explicit per-row column intervals, no query profile, no asm.
 */
#include "do_oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

enum { DIR_ALIGN = 1, DIR_INSERT = 2, DIR_DELETE = 4 }; /* EXT/alignmentMatrices.h:35-40 */

/* ------------------------------------------------------------------------------------------
 * Dense TCM  (EXT/c_code_alloc_setup.c:25-43 `distance`, :110-205 `setUp2dCostMtx`)
 * ---------------------------------------------------------------------------------------- */

/* min over the symbols `pos` present in ambiguity set `amb` of sigma[pos][nuc]  (nuc 0-based).
 * NB this indexes sigma[pos*a + nuc], the transpose of the Haskell overlap — identical only
 * for symmetric TCMs (reference comment "Requires symmetric"). */
static int tcm_distance(const uint32_t *sigma, int a, int nuc, uint32_t amb)
{
    int best = INT_MAX;
    for (int pos = 0; pos < a; pos++)
        if (amb & (1u << pos)) {
            int c = (int)sigma[pos * a + nuc];
            if (c < best) best = c;
        }
    return best;
}

oracle_tcm_t *oracle_tcm_new(const uint32_t *sigma, int a, uint32_t gap_open)
{
    oracle_tcm_t *t = calloc(1, sizeof *t);
    if (!t) return NULL;
    t->a = a;
    t->dim = 1u << a;
    t->gap = 1u << (a - 1);
    t->gap_open = gap_open;
    t->cost = calloc((size_t)t->dim * t->dim, sizeof(uint32_t));   /* row / column 0 stay zero */
    t->median = calloc((size_t)t->dim * t->dim, sizeof(uint32_t));
    if (!t->cost || !t->median) { oracle_tcm_free(t); return NULL; }
    /* per-element, per-symbol distances once, then combine */
    int *d = malloc(sizeof(int) * t->dim * (size_t)a);
    for (uint32_t e = 1; e < t->dim; e++)
        for (int s = 0; s < a; s++) d[e * a + s] = tcm_distance(sigma, a, s, e);
    for (uint32_t e1 = 1; e1 < t->dim; e1++)
        for (uint32_t e2 = 1; e2 < t->dim; e2++) {
            int best = INT_MAX;
            uint32_t med = 0;
            for (int s = 0; s < a; s++) {
                int c = d[e1 * a + s] + d[e2 * a + s];
                if (c < best)       { best = c; med = 1u << s; }
                else if (c == best) { med |= 1u << s; }
            }
            t->cost[((size_t)e1 << a) + e2] = (uint32_t)best;
            t->median[((size_t)e1 << a) + e2] = med;
        }
    free(d);
    return t;
}

void oracle_tcm_free(oracle_tcm_t *t)
{
    if (!t) return;
    free(t->cost);
    free(t->median);
    free(t);
}

static inline uint32_t C(const oracle_tcm_t *t, uint32_t x, uint32_t y)
{
    return t->cost[((size_t)x << t->a) + y];
}

/* ------------------------------------------------------------------------------------------
 * Band geometry  (EXT/c_alignment_interface.c:101-109; EXT/alignCharacters.c:1163-1374)
 * ---------------------------------------------------------------------------------------- */
uint64_t oracle_band_rows(size_t lg, size_t sh, int *mode, int32_t *lo, int32_t *hi, uint8_t *right_ukk)
{
    int diff = (int)lg - (int)sh;
    int lower_limit = (int)(.1 * (double)lg);
    int deltawh = diff < lower_limit ? lower_limit / 2 : 2;
    size_t W = 50 + (size_t)deltawh;
    size_t H = (lg - sh) + 50 + (size_t)deltawh;
    if (W > sh) W = sh;
    if (H > lg) H = lg;

    int m;
    if ((double)lg >= 1.5 * (double)sh) m = 0;
    else if (2 * H < lg)                 m = 2;
    else if (8 >= lg - H)                m = 0;
    else                                 m = 3;
    if (mode) *mode = m;

    uint64_t cells = 0;
    for (size_t i = 0; i < lg; i++) {
        int32_t l = 0, h = (int32_t)sh - 1;
        uint8_t ru = 0;
        if (m == 2) {
            if (i == 0)              { h = (int32_t)W - 1; }
            else if (i < H)          { h = (int32_t)(W + i) - 1; ru = 1; }          /* extending right      :675-746 */
            else if (i < lg - (H-1)) { l = (int32_t)(i - H) + 1; h = (int32_t)(i + W) - 1; ru = 1; } /* left+right :749-824 */
            else                     { l = (int32_t)sh - (int32_t)(W + H - 2) + (int32_t)(i - (lg - (H - 1))); } /* :827-920 */
        } else if (m == 3) {
            size_t k = sh - W;                                                       /* :1310-1370 */
            if (i == 0)              { h = (int32_t)W - 1; }
            else if (i < k + 1)      { h = (int32_t)(W + i) - 1; ru = 1; }
            else if (i < lg - k + 1) { /* full row */ }
            else                     { l = 1 + (int32_t)(i - (lg - k + 1)); }
        }
        if (lo) lo[i] = l;
        if (hi) hi[i] = h;
        if (right_ukk) right_ukk[i] = ru;
        cells += (uint64_t)(h - l + 1);
    }
    return cells;
}

/* ------------------------------------------------------------------------------------------
 * Fill + traceback + median  (EXT/alignCharacters.c:388-658, 4404-4507, 4569-4594)
 * lng / shr include the leading gap. Aligned outputs exclude the (gap,gap) column of cell (0,0).
 * ---------------------------------------------------------------------------------------- */
static int64_t linear_core(const oracle_tcm_t *t,
                           const uint32_t *lng, size_t lg, const uint32_t *shr, size_t sh,
                           uint32_t *al_lng, uint32_t *al_shr, size_t *out_len, uint64_t *cells_out)
{
    int32_t *lo = malloc(sizeof(int32_t) * lg), *hi = malloc(sizeof(int32_t) * lg);
    uint8_t *ru = malloc(lg);
    uint8_t *dir = malloc(lg * sh);
    uint32_t *rows = malloc(sizeof(uint32_t) * 2 * sh);
    int mode;
    uint64_t cells = oracle_band_rows(lg, sh, &mode, lo, hi, ru);
    if (cells_out) *cells_out = cells;

    uint32_t *prev = rows, *cur = rows + sh;
    const uint32_t gap = t->gap;

    /* row 0 (:615-641 / :1046-1052) */
    prev[0] = 0;
    dir[0] = DIR_ALIGN;
    for (int32_t j = 1; j <= hi[0]; j++) {
        prev[j] = prev[j - 1] + C(t, gap, shr[j]);
        dir[j] = DIR_INSERT;
    }
    for (size_t i = 1; i < lg; i++) {
        const uint32_t li = lng[i];
        const uint32_t c = C(t, li, gap);
        uint8_t *drow = dir + i * sh;
        int32_t j = lo[i];
        if (j == 0) {                                   /* first cell: DELETE only (:577-590, :644-658) */
            cur[0] = prev[0] + c;
            drow[0] = DIR_DELETE;
        } else {                                        /* algn_fill_ukk_left_cell (:516-554) */
            uint32_t up = prev[j] + c, dg = prev[j - 1] + C(t, li, shr[j]);
            if ((int)up < (int)dg)      { cur[j] = up; drow[j] = DIR_DELETE; }
            else if ((int)dg < (int)up) { cur[j] = dg; drow[j] = DIR_ALIGN; }
            else                        { cur[j] = up; drow[j] = DIR_ALIGN | DIR_DELETE; }
        }
        int32_t last3 = hi[i] - (ru[i] ? 1 : 0);
        for (j = j + 1; j <= last3; j++) {              /* algn_fill_row (:388-465) */
            uint32_t up = prev[j] + c;
            uint32_t lf = cur[j - 1] + C(t, gap, shr[j]);
            uint32_t dg = prev[j - 1] + C(t, li, shr[j]);
            uint32_t mn = up < lf ? up : lf;
            if (dg < mn) mn = dg;
            cur[j] = mn;
            drow[j] = (uint8_t)((dg == mn ? DIR_ALIGN : 0) | (lf == mn ? DIR_INSERT : 0) | (up == mn ? DIR_DELETE : 0));
        }
        if (ru[i] && hi[i] > lo[i]) {                   /* algn_fill_ukk_right_cell (:469-513) */
            j = hi[i];
            uint32_t lf = cur[j - 1] + C(t, gap, shr[j]), dg = prev[j - 1] + C(t, li, shr[j]);
            if ((int)lf < (int)dg)      { cur[j] = lf; drow[j] = DIR_INSERT; }
            else if ((int)dg < (int)lf) { cur[j] = dg; drow[j] = DIR_ALIGN; }
            else                        { cur[j] = dg; drow[j] = DIR_INSERT | DIR_ALIGN; }
        }
        uint32_t *tmp = prev; prev = cur; cur = tmp;
    }
    int64_t cost = (int64_t)prev[sh - 1];

    /* traceback, priority ALIGN > DELETE > INSERT (:4483-4507) */
    size_t i = lg - 1, j = sh - 1, n = 0, cap = lg + sh;
    uint32_t *rl = malloc(sizeof(uint32_t) * cap), *rs = malloc(sizeof(uint32_t) * cap);
    while (i > 0 || j > 0) {
        uint8_t d = dir[i * sh + j];
        if (d & DIR_ALIGN)       { rl[n] = lng[i]; rs[n] = shr[j]; i--; j--; }
        else if (d & DIR_DELETE) { rl[n] = lng[i]; rs[n] = 0;      i--;      }
        else                     { rl[n] = 0;      rs[n] = shr[j];      j--; }
        n++;
    }
    for (size_t k = 0; k < n; k++) { al_lng[k] = rl[n - 1 - k]; al_shr[k] = rs[n - 1 - k]; }
    *out_len = n;
    free(rl); free(rs); free(lo); free(hi); free(ru); free(dir); free(rows);
    return cost;
}

int64_t oracle_align2d(const oracle_tcm_t *t,
                       const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                       uint32_t *out_gapped, uint32_t *out_slot1, uint32_t *out_slot2,
                       size_t *out_len, uint64_t *cells)
{
    /* which input is "long": length, then first differing element (EXT/c_alignment_interface.c:51-85) */
    int first_long = 1;
    if (n1 < n2) first_long = 0;
    else if (n1 == n2)
        for (size_t i = 0; i < n1; i++)
            if (in1[i] != in2[i]) { first_long = in1[i] > in2[i]; break; }
    const uint32_t *pl = first_long ? in1 : in2, *ps = first_long ? in2 : in1;
    size_t lg = (first_long ? n1 : n2) + 1, sh = (first_long ? n2 : n1) + 1;

    uint32_t *lng = malloc(sizeof(uint32_t) * lg), *shr = malloc(sizeof(uint32_t) * sh);
    lng[0] = shr[0] = t->gap;                               /* alignIOtoDynChar (:461-478) */
    memcpy(lng + 1, pl, sizeof(uint32_t) * (lg - 1));
    memcpy(shr + 1, ps, sizeof(uint32_t) * (sh - 1));

    uint32_t *al_l = malloc(sizeof(uint32_t) * (lg + sh)), *al_s = malloc(sizeof(uint32_t) * (lg + sh));
    size_t n = 0;
    int64_t cost = linear_core(t, lng, lg, shr, sh, al_l, al_s, &n, cells);

    for (size_t k = 0; k < n; k++) {
        uint32_t x = al_s[k] ? al_s[k] : t->gap, y = al_l[k] ? al_l[k] : t->gap;
        out_gapped[k] = t->median[((size_t)x << t->a) + y];  /* (:4583-4592), arguments crossed at :115,:136 */
        /* the LONG input's buffer receives the aligned SHORT string and vice versa (:138-139) */
        out_slot1[k] = first_long ? al_s[k] : al_l[k];
        out_slot2[k] = first_long ? al_l[k] : al_s[k];
    }
    *out_len = n;
    free(lng); free(shr); free(al_l); free(al_s);
    return cost;
}

/* ------------------------------------------------------------------------------------------
 * Haskell-side wrapper
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t *m, *l, *r; size_t n; } dc_t;

/* DYN/Internal.hs:135-185 — strip triples whose median == gap; g[k] = gaps before ungapped index k
 * (trailing run at k == n). */
static void delete_gaps(const dc_t *c, uint32_t gap, dc_t *u, uint32_t *g)
{
    size_t n = 0;
    uint32_t run = 0;
    for (size_t i = 0; i < c->n; i++) {
        if (c->m[i] == gap) { run++; continue; }
        g[n] = run; run = 0;
        u->m[n] = c->m[i]; u->l[n] = c->l[i]; u->r[n] = c->r[i];
        n++;
    }
    g[n] = run;
    u->n = n;
}

/* DO/Pairwise/Internal.hs:307-337 — length, then medians, then (m,l,r) triples, lexicographic */
static int measure(const dc_t *x, const dc_t *y)
{
    if (x->n != y->n) return x->n < y->n ? -1 : 1;
    for (size_t i = 0; i < x->n; i++)
        if (x->m[i] != y->m[i]) return x->m[i] < y->m[i] ? -1 : 1;
    for (size_t i = 0; i < x->n; i++) {
        if (x->m[i] != y->m[i]) return x->m[i] < y->m[i] ? -1 : 1;
        if (x->l[i] != y->l[i]) return x->l[i] < y->l[i] ? -1 : 1;
        if (x->r[i] != y->r[i]) return x->r[i] < y->r[i] ? -1 : 1;
    }
    return 0;
}

int64_t oracle_foreign_pairwise_do(const oracle_tcm_t *t,
                                   const uint32_t *m1, const uint32_t *l1, const uint32_t *r1, size_t n1, int missing1,
                                   const uint32_t *m2, const uint32_t *l2, const uint32_t *r2, size_t n2, int missing2,
                                   uint32_t *om, uint32_t *ol, uint32_t *orr, size_t *on, int *omissing,
                                   uint64_t *cells)
{
    if (cells) *cells = 0;
    /* handleMissingCharacter (DO/Pairwise/Internal.hs:247-259) */
    if (missing1 || missing2) {
        int take2 = missing1 && !missing2;
        const uint32_t *sm = take2 ? m2 : m1, *sl = take2 ? l2 : l1, *sr = take2 ? r2 : r1;
        size_t sn = take2 ? n2 : n1;
        *omissing = take2 ? missing2 : missing1;
        *on = *omissing ? 0 : sn;
        for (size_t i = 0; i < *on; i++) { om[i] = sm[i]; ol[i] = sl[i]; orr[i] = sr[i]; }
        return 0;
    }
    const uint32_t gap = t->gap;
    dc_t c1 = { (uint32_t *)m1, (uint32_t *)l1, (uint32_t *)r1, n1 };
    dc_t c2 = { (uint32_t *)m2, (uint32_t *)l2, (uint32_t *)r2, n2 };
    size_t cap = n1 + n2 + 2;
    uint32_t *buf = malloc(sizeof(uint32_t) * cap * 11);
    dc_t u1 = { buf, buf + cap, buf + 2 * cap, 0 }, u2 = { buf + 3 * cap, buf + 4 * cap, buf + 5 * cap, 0 };
    uint32_t *g1 = buf + 6 * cap, *g2 = buf + 7 * cap;
    uint32_t *am = buf + 8 * cap, *al = buf + 9 * cap, *ar = buf + 10 * cap;
    delete_gaps(&c1, gap, &u1, g1);
    delete_gaps(&c2, gap, &u2, g2);

    /* measureAndUngapCharacters (DO/Pairwise/Internal.hs:364-381) */
    int ord = measure(&u1, &u2);
    if (ord == 0) ord = measure(&c1, &c2);
    int swapped = ord > 0;
    const dc_t *lesser = swapped ? &u2 : &u1, *longer = swapped ? &u1 : &u2;
    uint32_t *gl = swapped ? g2 : g1, *gr = swapped ? g1 : g2;

    int64_t cost = 0;
    size_t na = 0;
    if (lesser->n == 0) {                                   /* DO/Pairwise/FFI.hsc:240-248 */
        for (size_t i = 0; i < longer->n; i++) {
            uint32_t m = longer->m[i];
            am[i] = t->median[((size_t)m << t->a) + gap];
            al[i] = 0;
            ar[i] = m;
        }
        na = longer->n;
    } else {                                                /* longer is sent FIRST (FFI.hsc:261-268) */
        cost = oracle_align2d(t, longer->m, longer->n, lesser->m, lesser->n, am, al, ar, &na, cells);
    }

    /* insertGaps gapsLesser gapsLonger (DYN/Internal.hs:189-241) */
    size_t lp = 0, rp = 0, k = 0, o = 0;
    size_t total = na;
    for (size_t i = 0; i <= lesser->n; i++) total += gl[i];
    for (size_t i = 0; i <= longer->n; i++) total += gr[i];
    int any_gaps = total != na;
    for (o = 0; o < total; o++) {
        if (lp <= lesser->n && gl[lp] > 0)      { gl[lp]--; om[o] = gap; ol[o] = gap; orr[o] = 0; }
        else if (rp <= longer->n && gr[rp] > 0) { gr[rp]--; om[o] = gap; ol[o] = 0; orr[o] = gap; }
        else {
            om[o] = am[k]; ol[o] = al[k]; orr[o] = ar[k];
            if (al[k]) lp++;
            if (ar[k]) rp++;
            k++;
        }
    }
    if (swapped)                                            /* omap swapContext */
        for (o = 0; o < total; o++) { uint32_t x = ol[o]; ol[o] = orr[o]; orr[o] = x; }
    *on = total;
    /* both empty after ungapping and nothing to re-insert -> toMissing char1 */
    *omissing = (lesser->n == 0 && longer->n == 0 && !any_gaps);
    free(buf);
    return cost;
}
