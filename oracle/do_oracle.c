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

/* ------------------------------------------------------------------------------------------
 * C-affine dialect (see do_oracle.h).  Rows i = 1..S follow the SHORTER string, columns j = 1..L the
 * longer one; four cost planes (close-block-diagonal CB, extend-block-diagonal EB, extend-vertical EV,
 * extend-horizontal EH) kept as two rolling rows each, exactly like the reference, so that cells outside
 * the band carry whatever the reference would read there.
 * ---------------------------------------------------------------------------------------- */
enum {
    A2A = 1, A2V = 2, A2H = 4, A2D = 8, BEG_B = 16, END_B = 32, BEG_V = 64, END_V = 128, BEG_H = 256,
    END_H = 512, DO_A = 1024, DO_V = 2048, DO_H = 4096, DO_D = 8192          /* alignCharacters.c:76-89 */
};
#define AFF_INF 100000u                                                       /* alignCharacters.h:70 */

static uint32_t aff_sub(const oracle_tcm_t *t, int variant, uint32_t x, uint32_t y)
{
    if (variant) return t->cost[((size_t)x << t->a) + y];
    /* as compiled for x86-64: the 32-bit shift count (costMatrixDimension = 2^a) is taken mod 32 */
    return t->cost[((size_t)x << (t->dim & 31u)) + y];
}

int64_t oracle_align2d_affine(const oracle_tcm_t *t, int variant,
                              const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                              uint32_t *out_gapped, uint32_t *out_slot1, uint32_t *out_slot2,
                              size_t *out_len, uint64_t *cells_out)
{
    const int first_long = n1 >= n2;                                      /* c_alignment_interface.c:212 */
    const size_t L = first_long ? n1 : n2, S = first_long ? n2 : n1;
    const uint32_t g = t->gap, GO = t->gap_open, amb = g - 1;
    uint32_t *lng = malloc(sizeof(uint32_t) * (L + 1)), *shr = malloc(sizeof(uint32_t) * (S + 1));
    lng[0] = shr[0] = g;
    memcpy(lng + 1, first_long ? in1 : in2, sizeof(uint32_t) * L);
    memcpy(shr + 1, first_long ? in2 : in1, sizeof(uint32_t) * S);

    const size_t W = L + 1;
    uint32_t *buf = calloc(8 * W + 3 * W, sizeof(uint32_t));
    uint32_t *CB[2] = { buf, buf + W }, *EB[2] = { buf + 2 * W, buf + 3 * W };
    uint32_t *EV[2] = { buf + 4 * W, buf + 5 * W }, *EH[2] = { buf + 6 * W, buf + 7 * W };
    uint32_t *gr = buf + 8 * W, *go = buf + 9 * W, *he = buf + 10 * W;
    uint16_t *dir = calloc((S + 1) * W, sizeof(uint16_t));
    uint32_t fc_last = 0;
    uint64_t cells = 0;

    /* row 0 (alignCharacters.c:3152-3173) */
    CB[0][0] = 0; EB[0][0] = 0; EH[0][0] = GO; EV[0][0] = GO; dir[0] = 0xFFFF;
    for (size_t j = 1; j <= L; j++) {
        gr[j] = C(t, g, lng[j]);
        const uint32_t r = EH[0][j - 1] + gr[j];
        EH[0][j] = r; CB[0][j] = r; EB[0][j] = AFF_INF; EV[0][j] = AFF_INF;
        dir[j] = DO_H | END_H;
        if (j == L) fc_last = r;
    }
    /* per-column gap-open / horizontal-extension costs (:3585-3595) */
    for (size_t j = 1; j <= L; j++) {
        go[j] = (!(g & lng[j - 1]) && (g & lng[j])) ? 0 : GO;
        he[j] = ((lng[j - 1] & g) && !(lng[j] & g)) ? go[j] + gr[j] : gr[j];
    }
    if (L >= 1) he[1] = gr[1];

    size_t start = 1, end = (L - S) + 8;
    if (end < 40) end = 40;
    if (end > L) end = L;
    for (size_t i = 1; i <= S; i++) {
        uint32_t *cb = CB[i & 1], *eb = EB[i & 1], *ev = EV[i & 1], *eh = EH[i & 1];
        const uint32_t *pcb = CB[(i - 1) & 1], *peb = EB[(i - 1) & 1], *pev = EV[(i - 1) & 1], *peh = EH[(i - 1) & 1];
        uint16_t *d = dir + i * W;
        if (i > 40) start++;
        const uint32_t si = shr[i], sp = shr[i - 1], sng = si & amb;
        const uint32_t ge = C(t, si, g);
        const uint32_t so = (!(g & sp) && (g & si)) ? 0 : GO;
        const uint32_t ve = (i > 1 && (sp & g) && !(si & g)) ? so + ge : ge;
        /* border cell: vertical extension only (:3614-3634) */
        const uint32_t r = pev[start - 1] + ve;
        eh[start - 1] = AFF_INF; eb[start - 1] = AFF_INF; ev[start - 1] = r; cb[start - 1] = AFF_INF;
        d[start - 1] = DO_V | END_V;
        if (start - 1 == L) fc_last = r;
        for (size_t j = start; j <= end; j++) {
            const uint32_t lj = lng[j], lng_ = lj & amb;
            uint16_t m = 0;
            /* FILL_EXTEND_HORIZONTAL (:2527-2560): signed compare, strict */
            int ext = (int)(eh[j - 1] + he[j]), opn = (int)(cb[j - 1] + go[j] + gr[j]);
            if (ext < opn) { m |= BEG_H; eh[j] = (uint32_t)ext; } else { m |= END_H; eh[j] = (uint32_t)opn; }
            /* FILL_EXTEND_VERTICAL (:2588-2617) */
            ext = (int)(pev[j] + ve); opn = (int)(pcb[j] + so + ge);
            if (ext < opn) { m |= BEG_V; ev[j] = (uint32_t)ext; } else { m |= END_V; ev[j] = (uint32_t)opn; }
            /* FILL_EXTEND_BLOCK_DIAGONAL (:2660-2710): both options add the same `diag` */
            const uint32_t dg0 = ((g & si) && (g & lj)) ? 0 : AFF_INF;
            ext = (int)(peb[j - 1] + dg0); opn = (int)(pcb[j - 1] + dg0);
            if (ext < opn) { m |= BEG_B; eb[j] = (uint32_t)ext; } else { m |= END_B; eb[j] = (uint32_t)opn; }
            /* FILL_CLOSE_BLOCK_DIAGONAL (:2770-2848): unsigned compares */
            const uint32_t dg = aff_sub(t, variant, sng, lng_);
            const uint32_t xgo = go[j] < so ? so : go[j];
            const uint32_t fv = (si == sng) ? pev[j - 1] + dg : pev[j - 1] + dg + go[j];
            const uint32_t fh = (lj == lng_) ? peh[j - 1] + dg : peh[j - 1] + dg + so;
            const uint32_t fd = peb[j - 1] + dg + xgo;
            uint32_t c = pcb[j - 1] + dg;
            uint16_t mk = A2A;
            if (c >= fv) { if (c > fv) { c = fv; mk = A2V; } else mk |= A2V; }
            if (c >= fh) { if (c > fh) { c = fh; mk = A2H; } else mk |= A2H; }
            if (c >= fd) { if (c > fd) { c = fd; mk = A2D; } else mk |= A2D; }
            cb[j] = c;
            m |= mk;
            /* ASSIGN_MINIMUM (:3218-3256) */
            uint32_t f = eh[j];
            mk = DO_H;
            if (f >= ev[j]) { if (f > ev[j]) { f = ev[j]; mk = DO_V; } else mk |= DO_V; }
            if (f >= eb[j]) { if (f > eb[j]) { f = eb[j]; mk = DO_D; } else mk |= DO_D; }
            if (f >= cb[j]) { if (f > cb[j]) { f = cb[j]; mk = DO_A; } else mk |= DO_A; }
            m |= mk;
            d[j] = m;
            if (j == L) fc_last = f;
            cells++;
        }
        if (end < L) {
            end++;
            d[end] = DO_H | END_H;
            ev[end] = AFF_INF; cb[end] = AFF_INF; eh[end] = AFF_INF; eb[end] = AFF_INF;
        }
    }
    if (cells_out) *cells_out = cells;
    const int64_t cost = (int64_t)fc_last;

    /* traceback state machine (:2850-3010) */
    enum { M_TODO, M_V, M_H, M_D, M_A } mode = M_TODO;
    size_t si_ = S, li_ = L, n = 0, cap = S + L + 2;
    uint32_t *rm = malloc(sizeof(uint32_t) * cap), *rs = malloc(sizeof(uint32_t) * cap), *rl = malloc(sizeof(uint32_t) * cap);
    while (si_ != 0 && li_ != 0) {
        const uint16_t fl = dir[si_ * W + li_];
        const uint32_t se = shr[si_], le = lng[li_];
        if (mode == M_TODO) {
            if (fl & DO_H) mode = M_H; else if (fl & DO_A) mode = M_A; else if (fl & DO_V) mode = M_V; else mode = M_D;
        } else if (mode == M_V) {
            if (fl & END_V) mode = M_TODO;
            rm[n] = !(se & g) ? (se | g) : g; rs[n] = se; rl[n] = g; n++; si_--;
        } else if (mode == M_H) {
            if (fl & END_H) mode = M_TODO;
            rm[n] = !(le & g) ? (le | g) : g; rs[n] = g; rl[n] = le; n++; li_--;
        } else if (mode == M_D) {
            if (fl & END_B) mode = M_TODO;
            rm[n] = g; rs[n] = se; rl[n] = le; n++; si_--; li_--;
        } else {
            if (fl & A2H) mode = M_H; else if (fl & A2D) mode = M_D; else if (fl & A2V) mode = M_V;
            rm[n] = t->median[((size_t)(se & amb) << t->a) + (le & amb)]; rs[n] = se; rl[n] = le; n++; si_--; li_--;
        }
    }
    while (si_ != 0) { const uint32_t se = shr[si_]; rm[n] = !(se & g) ? (se | g) : g; rs[n] = se; rl[n] = g; n++; si_--; }
    while (li_ != 0) { const uint32_t le = lng[li_]; rm[n] = !(le & g) ? (le | g) : g; rs[n] = g; rl[n] = le; n++; li_--; }
    for (size_t k = 0; k < n; k++) {
        out_gapped[k] = rm[n - 1 - k];
        const uint32_t as = rs[n - 1 - k], al = rl[n - 1 - k];
        out_slot1[k] = first_long ? al : as;            /* natural order: each slot keeps its own string */
        out_slot2[k] = first_long ? as : al;
    }
    *out_len = n;
    free(rm); free(rs); free(rl); free(dir); free(buf); free(lng); free(shr);
    return cost;
}

/* ------------------------------------------------------------------------------------------
 * Haskell dialects S1 / S2 over an arbitrary overlap function (alphabets up to 32 symbols here).
 * PARITY UNPINNED: no GHC in this environment, no golden vector upstream for these paths; restated from
 *   DO/Pairwise/Internal.hs:389-424 (needlemanWunschDefinition, S1 cell), NeedlemanWunsch.hs:97-143,
 *   UnboxedSwapping.hs:58-124 (S1 full), Ukkonen/Internal.hs:79-202 + Ukkonen/Ribbon.hs:100-155 (S1 ribbon),
 *   UnboxedUkkonenFullSpace.hs:59-276, 641-815 (S2 band: getMinimalResult relabels Diag -> Left when the
 *   median is the gap), UnboxedUkkonenSwapping.hs:436-459 (ukkonenConstants),
 *   Bio/Metadata/Overlap.hs:56-86 (overlap), Data/MetricRepresentation.hs:116-128 (discrete metric).
 * The Ukkonen variants are filled FROM SCRATCH at every offset of the reference's doubling schedule (the
 * reference expands its band incrementally, UnboxedUkkonenFullSpace.hs:344-638; equal results assumed).
 * Rows i = 0..m follow the lesser string, columns j = 0..n the longer one, exactly as in the Haskell.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int kind;                /* 0 discrete metric, 1 explicit symbol-change matrix, 2 L1 norm (closed form), 9 external lookup (compiled tcm-memo) */
    int a;
    const uint32_t *sigma;   /* a*a, indexed [median symbol][input symbol] (Overlap.hs:75-86) */
    oracle_lookup_fn lookup; /* kind 9 */
    void *lookup_ctx;
} hs_metric_t;

static void hs_overlap(const hs_metric_t *mt, uint32_t x, uint32_t y, uint32_t *med, uint64_t *cost)
{
    if (mt->kind == 0) {                                     /* discreteMetricPairwiseLogic */
        if (x & y) { *med = x & y; *cost = 0; } else { *med = x | y; *cost = 1; }
        return;
    }
    if (mt->kind == 9) { *cost = mt->lookup(mt->lookup_ctx, x, y, med); return; }
    if (mt->kind == 2) {
        /* firstLinearNormPairwiseLogic (Data/MetricRepresentation.hs:166-190) over Ranged AmbiguityGroup
         * (Bio/Character/Encodable/Dynamic/AmbiguityGroup.hs:154-172) and Data/Range.hs:100-190.  toRange = (lowest set
         * symbol, highest set symbol) — PINNED by the L1 script costs 3483 / 21851 through hs_wide.c, the same logic on
         * multi-word elements; fromRange of lb < ub sets symbols lb .. ub-1 only ((2^ub - 1) xor (2^lb - 1)). */
        const int a = mt->a;
        const int xl = x ? __builtin_ctz(x) : a, xh = x ? 31 - __builtin_clz(x) : 0;
        const int yl = y ? __builtin_ctz(y) : a, yh = y ? 31 - __builtin_clz(y) : 0;
        const int lo0 = xl < xh ? xl : xh, hi0 = xl < xh ? xh : xl, lo1 = yl < yh ? yl : yh, hi1 = yl < yh ? yh : yl;
        const int isect = lo1 <= hi0 && hi1 >= lo0;
        int nl, nu;
        if (isect) { nl = lo0 > lo1 ? lo0 : lo1; nu = hi0 < hi1 ? hi0 : hi1; }
        else       { nl = hi0 < hi1 ? hi0 : hi1; nu = lo0 > lo1 ? lo0 : lo1; }
        const uint64_t full = (a >= 32) ? 0xffffffffull : ((1ull << a) - 1);
        *med = (uint32_t)((nu == nl ? (1ull << nu) : (((1ull << nu) - 1) ^ ((1ull << nl) - 1))) & full);
        *cost = isect ? 0 : (uint64_t)(nu - nl);
        return;
    }
    uint64_t best = UINT64_MAX;
    uint32_t bits = 0;
    for (int s = mt->a - 1; s >= 0; s--) {                   /* overlap: every symbol attaining the minimum */
        uint64_t dx = UINT64_MAX, dy = UINT64_MAX;
        for (int j = 0; j < mt->a; j++) {
            if (x & (1u << j)) { uint64_t v = mt->sigma[s * mt->a + j]; if (v < dx) dx = v; }
            if (y & (1u << j)) { uint64_t v = mt->sigma[s * mt->a + j]; if (v < dy) dy = v; }
        }
        const uint64_t c = dx + dy;
        if (c == best) bits |= 1u << s;
        else if (c < best) { best = c; bits = 1u << s; }
    }
    *med = bits; *cost = best;
}

/* exported: one overlap lookup (metric_kind as oracle_haskell_do); the explicit branch restates
 * CostMatrix_2d::computeCostMedian / findDistance (lib/tcm-memo/ffi/memoized-tcm/costMatrix_2d.cpp:175-267) and is
 * PINNED against the compiled reference tcm-memo (tests/golden/tcm_memo_golden.json, tests/test_oracle.py). */
uint64_t oracle_overlap(int metric_kind, const uint32_t *sigma, int a, uint32_t x, uint32_t y, uint32_t *med)
{
    const hs_metric_t mt = { metric_kind, a, sigma, NULL, NULL };
    uint64_t c;
    hs_overlap(&mt, x, y, med, &c);
    return c;
}

#define HS_INF (UINT64_MAX / 4)
enum { HS_DIAG = 0, HS_LEFT = 1, HS_UP = 2 };

/* one banded fill: cells with j - i outside [-off, (n-m) + off] are infinite. s2 != 0: getMinimalResult. */
static uint64_t hs_fill(const hs_metric_t *mt, int s2, const uint32_t *les, size_t m, const uint32_t *lon, size_t n,
                        long off, uint8_t *dir, uint64_t *cells)
{
    const uint32_t gap = 1u << (mt->a - 1);
    const size_t W = n + 1;
    uint64_t *prev = malloc(sizeof(uint64_t) * W), *cur = malloc(sizeof(uint64_t) * W), cnt = 0;
    /* s2 bit 0: getMinimalResult's relabel; bit 1: unboxedFullMatrixDO's first column, seeded from ROW 0 —
     * (i,0) = cost(left_i, gap) + m(0, i-1), "firstPrevCost <- ref (0, i - 1)" (UnboxedFullMatrix.hs:86-90), restated
     * literally (rows <= columns, so the cell exists) */
    const int quirk = (s2 & 2) != 0;
    uint64_t *row0 = quirk ? malloc(sizeof(uint64_t) * W) : NULL;
    s2 &= 1;
    for (size_t i = 0; i <= m; i++) {
        for (size_t j = 0; j <= n; j++) {
            const long x = (long)j - (long)i;
            if (x < -off || x > (long)(n - m) + off) { cur[j] = HS_INF; continue; }
            cnt++;
            if (i == 0 && j == 0) { cur[0] = 0; dir[0] = HS_DIAG; if (quirk) row0[0] = 0; continue; }
            uint32_t md; uint64_t c;
            if (quirk && j == 0) {
                hs_overlap(mt, les[i - 1], gap, &md, &c);
                cur[0] = row0[i - 1] + c;
                dir[i * W] = HS_UP;
                continue;
            }
            uint64_t best = HS_INF; uint8_t bd = HS_DIAG; int relabel = 0;
            if (i > 0 && j > 0 && prev[j - 1] < HS_INF) {
                hs_overlap(mt, lon[j - 1], les[i - 1], &md, &c);
                best = prev[j - 1] + c; bd = HS_DIAG; relabel = s2 && md == gap;
            }
            if (j > 0 && cur[j - 1] < HS_INF) {
                hs_overlap(mt, lon[j - 1], gap, &md, &c);
                if (cur[j - 1] + c < best) { best = cur[j - 1] + c; bd = HS_LEFT; relabel = 0; }
            }
            if (i > 0 && prev[j] < HS_INF) {
                hs_overlap(mt, gap, les[i - 1], &md, &c);
                if (prev[j] + c < best) { best = prev[j] + c; bd = HS_UP; relabel = 0; }
            }
            cur[j] = best;
            if (quirk && i == 0) row0[j] = best;
            dir[i * W + j] = relabel ? HS_LEFT : bd;
        }
        uint64_t *t = prev; prev = cur; cur = t;
    }
    const uint64_t res = prev[n];
    free(prev); free(cur); free(row0);
    if (cells) *cells += cnt;
    return res;
}

/* dialect: 0 S1 full (naiveDO / naiveDOMemo / unboxedSwappingDO), 1 S1 Ukkonen (ukkonenDO),
 *          2 S2 Ukkonen (unboxedUkkonenFullSpaceDO), 3 S2 full with the first column seeded from row 0
 *          (unboxedFullMatrixDO, as coded).  in1 / in2: median strings (no gap-only elements).
 * Outputs (length *out_len): median, aligned first input, aligned second input (0 = gap). */
static int64_t hs_do(const hs_metric_t mt, int dialect, int a,
                     const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                     uint32_t *out_med, uint32_t *out_a1, uint32_t *out_a2, size_t *out_len,
                     uint64_t *cells, int *final_offset);

int64_t oracle_haskell_do(int dialect, int metric_kind, const uint32_t *sigma, int a,
                          const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                          uint32_t *out_med, uint32_t *out_a1, uint32_t *out_a2, size_t *out_len,
                          uint64_t *cells, int *final_offset)
{
    const hs_metric_t mt = { metric_kind, a, sigma, NULL, NULL };
    return hs_do(mt, dialect, a, in1, n1, in2, n2, out_med, out_a1, out_a2, out_len, cells, final_offset);
}

/* the same dialects with every overlap answered by an external lookup (bench: the compiled reference tcm-memo) */
int64_t oracle_haskell_do_lookup(int dialect, oracle_lookup_fn lookup, void *lookup_ctx, int a,
                                 const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                                 uint32_t *out_med, uint32_t *out_a1, uint32_t *out_a2, size_t *out_len,
                                 uint64_t *cells, int *final_offset)
{
    const hs_metric_t mt = { 9, a, NULL, lookup, lookup_ctx };
    return hs_do(mt, dialect, a, in1, n1, in2, n2, out_med, out_a1, out_a2, out_len, cells, final_offset);
}

static int64_t hs_do(const hs_metric_t mt, int dialect, int a,
                     const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                     uint32_t *out_med, uint32_t *out_a1, uint32_t *out_a2, size_t *out_len,
                     uint64_t *cells, int *final_offset)
{
    const uint32_t gap = 1u << (a - 1);
    /* measureCharacters on median strings: shorter first, then lexicographic; equal -> not swapped */
    int swapped = 0;
    if (n1 != n2) swapped = n1 > n2;
    else for (size_t i = 0; i < n1; i++) if (in1[i] != in2[i]) { swapped = in1[i] > in2[i]; break; }
    const uint32_t *les = swapped ? in2 : in1, *lon = swapped ? in1 : in2;
    const size_t m = swapped ? n2 : n1, n = swapped ? n1 : n2;
    if (cells) *cells = 0;
    if (final_offset) *final_offset = -1;
    uint8_t *dir = calloc((m + 1) * (n + 1), 1);

    /* Delta = cheapest indel of a single non-gap symbol; G = elements containing the gap bit */
    uint64_t delta = UINT64_MAX, G = 0;
    for (int s = 0; s < a - 1; s++) {
        uint32_t md; uint64_t c1, c2;
        hs_overlap(&mt, 1u << s, gap, &md, &c1);
        hs_overlap(&mt, gap, 1u << s, &md, &c2);
        const uint64_t c = c1 < c2 ? c1 : c2;
        if (c < delta) delta = c;
    }
    for (size_t i = 0; i < m; i++) G += (les[i] & gap) != 0;
    for (size_t j = 0; j < n; j++) G += (lon[j] & gap) != 0;
    const uint64_t qd = n - m + 1;
    int no_gain = m <= 4 || 2 * n >= 3 * m || delta == 0;
    if (dialect == 2) no_gain = no_gain || G >= n;

    uint64_t cost;
    if (dialect == 3) {                                         /* unboxedFullMatrixDO (UnboxedFullMatrix.hs:57-144) */
        cost = hs_fill(&mt, 3, les, m, lon, n, (long)(m + n + 2), dir, cells);   /* full matrix, S2 cell, row-0-seeded column */
    } else if (dialect == 0 || no_gain) {
        cost = hs_fill(&mt, 0, les, m, lon, n, (long)(m + n + 2), dir, cells);   /* full matrix, S1 cell */
    } else if (dialect == 1) {                                  /* Ukkonen/Internal.hs:130-202 */
        uint64_t off = 2;
        for (;;) {
            const long a_off = (long)(off < m ? off : m);       /* Ribbon.hs: a = min alpha (w - d) */
            cost = hs_fill(&mt, 0, les, m, lon, n, a_off, dir, cells);
            const int64_t cv = (int64_t)delta * ((int64_t)qd + (int64_t)off - (int64_t)G);
            const uint64_t thr = cv > 0 ? (uint64_t)cv : 0;
            if (final_offset) *final_offset = (int)off;
            if (thr <= cost) off *= 2; else break;
        }
    } else {                                                    /* UnboxedUkkonenFullSpace.hs:159-207 */
        uint64_t off = 2 + G;
        for (;;) {
            const long a_off = (long)(off < m ? off : m);       /* ukkonenConstants: min o (cols - qd) */
            cost = hs_fill(&mt, 1, les, m, lon, n, a_off, dir, cells);
            if (final_offset) *final_offset = (int)off;
            if (qd + off > n) break;
            const uint64_t thr = (qd + off <= G) ? 0 : delta * (qd + off - G);
            if (thr <= cost) off *= 2; else break;
        }
    }
    /* traceback (Pairwise/Internal.hs:531-581) */
    size_t i = m, j = n, k = 0, cap = m + n + 1;
    uint32_t *rm = malloc(sizeof(uint32_t) * cap), *rx = malloc(sizeof(uint32_t) * cap), *ry = malloc(sizeof(uint32_t) * cap);
    while (i > 0 || j > 0) {
        const uint8_t d = dir[i * (n + 1) + j];
        uint32_t md; uint64_t c;
        if (d == HS_LEFT)    { hs_overlap(&mt, gap, lon[j - 1], &md, &c); rm[k] = md; rx[k] = 0; ry[k] = lon[j - 1]; j--; }
        else if (d == HS_UP) { hs_overlap(&mt, les[i - 1], gap, &md, &c); rm[k] = md; rx[k] = les[i - 1]; ry[k] = 0; i--; }
        else                 { hs_overlap(&mt, les[i - 1], lon[j - 1], &md, &c); rm[k] = md; rx[k] = les[i - 1]; ry[k] = lon[j - 1]; i--; j--; }
        k++;
    }
    for (size_t t = 0; t < k; t++) {
        out_med[t] = rm[k - 1 - t];
        const uint32_t x = rx[k - 1 - t], y = ry[k - 1 - t];     /* x from the lesser, y from the longer */
        out_a1[t] = swapped ? y : x;
        out_a2[t] = swapped ? x : y;
    }
    *out_len = k;
    free(rm); free(rx); free(ry); free(dir);
    return (int64_t)cost;
}
