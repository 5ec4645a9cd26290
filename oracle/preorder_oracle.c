/* preorder_oracle.c — CPU restatement of PCG's pre-order pass for dynamic characters (implied alignment).
 *
 * TEST INFRASTRUCTURE ONLY (see do_oracle.h).
 *
 * Restates, for one rooted binary tree and one dynamic character under a dense TCM:
 *   directOptimizationPreorder / initializeRoot / updateFromParent   DO/Internal.hs:148-233
 *   deriveImpliedAlignment                                           DO/Internal.hs:236-286
 *   lexicallyDisambiguate / disambiguateElement                      DO/Internal.hs:186-211
 *   getContext, gapElement                                           DYN/Element.hs:135-149
 * on top of the post-order contexts of oracle_foreign_pairwise_do (do_oracle.c).
 *
 * PARITY UNPINNED: upstream holds no golden vector for this pass — chel_xml.golden prints the alignment context in the
 * implied-alignment field (33 of 33 nodes identical) and the script tests that would exercise it are commented out
 * ("Impossible Happened in Implied Alignment").  The restatement is checked against an independent Python restatement
 * of the same functions (pcg_b200/preorder.py) and through the invariants of the pass (tests/test_oracle.py).
 */
#include <stdlib.h>
#include <string.h>

#include "do_oracle.h"

enum { CTX_GAPPING = 0, CTX_DELETION = 1, CTX_INSERTION = 2, CTX_ALIGNMENT = 3 };

/* DYN/Element.hs:144-149: (left non-empty, right non-empty) */
static int get_context(uint32_t l, uint32_t r)
{
    return (l != 0 && r != 0) ? CTX_ALIGNMENT : (l != 0) ? CTX_INSERTION : (r != 0) ? CTX_DELETION : CTX_GAPPING;
}

/* DO/Internal.hs:236-286.  One left fold over the parent's final alignment; the output has its length.
 * Returns 0, or -1 where the reference calls `error "Impossible happened in 'deriveImpliedAlignment'"`. */
int oracle_derive_implied_alignment(int is_left_child, uint32_t gap,
                                    const uint32_t *pam, const uint32_t *pal, const uint32_t *par, size_t n,
                                    const uint32_t *pcl, const uint32_t *pcr, size_t np,
                                    const uint32_t *ccm, const uint32_t *ccl, const uint32_t *ccr, size_t nc,
                                    uint32_t *om, uint32_t *ol, uint32_t *orr)
{
    size_t xi = 0, yi = 0;
    for (size_t k = 0; k < n; k++) {
        if (xi >= nc) {                                   /* go (acc, [], _) _ = (gap : acc, [], []) */
            om[k] = gap; ol[k] = 0; orr[k] = 0;
            continue;
        }
        if (yi >= np) return -1;                          /* go (_, _, []) _ = error ... */
        int take = 0, keep = 0;
        switch (get_context(pal[k], par[k])) {
            case CTX_GAPPING:   keep = 1; break;                                                          /* (e : acc, x:xs, y:ys) */
            case CTX_ALIGNMENT: take = 1; yi++; break;                                                    /* (x : acc, xs, ys)     */
            case CTX_DELETION:  take = get_context(pcl[yi], pcr[yi]) == CTX_DELETION && !is_left_child; yi++; break;
            default:            take = get_context(pcl[yi], pcr[yi]) == CTX_INSERTION && is_left_child; yi++; break;
        }
        if (keep)      { om[k] = pam[k]; ol[k] = pal[k]; orr[k] = par[k]; }
        else if (take) { om[k] = ccm[xi]; ol[k] = ccl[xi]; orr[k] = ccr[xi]; xi++; }
        else           { om[k] = gap; ol[k] = 0; orr[k] = 0; }                                            /* (gap : acc, x:xs, ys) */
    }
    return 0;
}

/* DO/Internal.hs:200-211: every element becomes (v, v, v), v = the single symbol at index
 * min(a - 1, countLeadingZeros median); bv-little counts from index 0 (settled by the L1 script costs),
 * so v is the LOWEST set symbol and an empty median gives the gap. */
void oracle_lexically_disambiguate(const uint32_t *med, size_t n, int a, uint32_t *out)
{
    for (size_t k = 0; k < n; k++) {
        uint32_t m = med[k] & ((a >= 32) ? 0xffffffffu : ((1u << a) - 1u));
        int idx = a - 1;
        if (m) { idx = 0; while (!((m >> idx) & 1u)) idx++; if (idx > a - 1) idx = a - 1; }
        out[k] = 1u << idx;
    }
}

/* Post-order contexts + pre-order implied alignments of one tree, one character.
 * Node ids: leaves 0 .. n_taxa-1, internal node x = n_taxa + x, children[2x], children[2x+1] (left, right).
 * Outputs per node v (2*n_taxa - 1 of them), `cap` elements each:
 *   ctx_*[v*cap ..]  post-order context (median, left, right), ctx_len[v]   (leaf: (s,s,s))
 *   ia_*[v*cap ..]   final (implied) alignment, ia_len[v];   single[v*cap ..] its lexical disambiguation
 * A length of 0xffffffff marks a Missing character.  Returns the root's node id, -1 on a bad tree / capacity, -2 on the
 * reference's "Impossible happened". */
int oracle_preorder_tree(const oracle_tcm_t *t, uint32_t n_taxa, const int32_t *children,
                         const uint32_t *const *leaf, const size_t *leaf_len, size_t cap,
                         uint32_t *ctx_m, uint32_t *ctx_l, uint32_t *ctx_r, uint32_t *ctx_len,
                         uint32_t *ia_m, uint32_t *ia_l, uint32_t *ia_r, uint32_t *ia_len, uint32_t *single)
{
    const uint32_t n_nodes = 2 * n_taxa - 1, n_int = n_taxa - 1, MISSING = 0xffffffffu;
    int *done = calloc(n_nodes, sizeof(int)), *parent = malloc(sizeof(int) * n_nodes);
    int rc = 0, root = -1;
    for (uint32_t v = 0; v < n_nodes; v++) parent[v] = -1;
    for (uint32_t x = 0; x < n_int; x++)
        for (int s = 0; s < 2; s++) {
            int c = children[2 * x + s];
            if (c < 0 || (uint32_t)c >= n_nodes || parent[c] != -1) { rc = -1; goto out; }
            parent[c] = (int)(n_taxa + x);
        }
    for (uint32_t v = 0; v < n_nodes; v++)
        if (parent[v] == -1) { if (root != -1) { rc = -1; goto out; } root = (int)v; }
    for (uint32_t v = 0; v < n_taxa; v++) {                 /* initializeLeaf (DO/Internal.hs:109-121): (s, s, s) */
        if (leaf_len[v] > cap) { rc = -1; goto out; }
        for (size_t i = 0; i < leaf_len[v]; i++) ctx_m[v * cap + i] = ctx_l[v * cap + i] = ctx_r[v * cap + i] = leaf[v][i];
        ctx_len[v] = (uint32_t)leaf_len[v];
        done[v] = 1;
    }
    for (uint32_t pass = 0, left = n_int; left && pass <= n_int; pass++)      /* post-order: children before parents */
        for (uint32_t x = 0; x < n_int; x++) {
            const uint32_t v = n_taxa + x;
            const int l = children[2 * x], r = children[2 * x + 1];
            if (done[v] || !done[l] || !done[r]) continue;
            const int ml = ctx_len[l] == MISSING, mr = ctx_len[r] == MISSING;
            const size_t nl = ml ? 0 : ctx_len[l], nr = mr ? 0 : ctx_len[r];
            if (nl + nr + 2 > cap) { rc = -1; goto out; }
            size_t on = 0;
            int omiss = 0;
            oracle_foreign_pairwise_do(t, ctx_m + l * cap, ctx_l + l * cap, ctx_r + l * cap, nl, ml,
                                       ctx_m + r * cap, ctx_l + r * cap, ctx_r + r * cap, nr, mr,
                                       ctx_m + v * cap, ctx_l + v * cap, ctx_r + v * cap, &on, &omiss, NULL);
            ctx_len[v] = omiss ? MISSING : (uint32_t)on;
            done[v] = 1;
            left--;
        }
    for (uint32_t v = 0; v < n_nodes; v++) if (!done[v]) { rc = -1; goto out; }

    /* pre-order: initializeRoot (:166-177), then updateFromParent (:216-233) down the tree */
    for (uint32_t v = 0; v < n_nodes; v++) done[v] = 0;
    ia_len[root] = ctx_len[root];
    if (ctx_len[root] != MISSING) {
        memcpy(ia_m + (size_t)root * cap, ctx_m + (size_t)root * cap, sizeof(uint32_t) * ctx_len[root]);
        memcpy(ia_l + (size_t)root * cap, ctx_l + (size_t)root * cap, sizeof(uint32_t) * ctx_len[root]);
        memcpy(ia_r + (size_t)root * cap, ctx_r + (size_t)root * cap, sizeof(uint32_t) * ctx_len[root]);
        oracle_lexically_disambiguate(ia_m + (size_t)root * cap, ctx_len[root], t->a, single + (size_t)root * cap);
    }
    done[root] = 1;
    for (uint32_t pass = 0; pass < n_nodes; pass++)
        for (uint32_t x = 0; x < n_int; x++) {
            const uint32_t p = n_taxa + x;
            if (!done[p]) continue;
            for (int s = 0; s < 2; s++) {
                const uint32_t c = (uint32_t)children[2 * x + s];
                if (done[c]) continue;
                if (ctx_len[c] == MISSING || ia_len[p] == MISSING) {          /* isMissing cac: inherit the parent's */
                    ia_len[c] = ia_len[p];
                    if (ia_len[p] != MISSING) {
                        memcpy(ia_m + (size_t)c * cap, ia_m + (size_t)p * cap, sizeof(uint32_t) * ia_len[p]);
                        memcpy(ia_l + (size_t)c * cap, ia_l + (size_t)p * cap, sizeof(uint32_t) * ia_len[p]);
                        memcpy(ia_r + (size_t)c * cap, ia_r + (size_t)p * cap, sizeof(uint32_t) * ia_len[p]);
                        memcpy(single + (size_t)c * cap, single + (size_t)p * cap, sizeof(uint32_t) * ia_len[p]);
                    }
                } else {
                    if (oracle_derive_implied_alignment(s == 0, t->gap, ia_m + (size_t)p * cap, ia_l + (size_t)p * cap,
                                                        ia_r + (size_t)p * cap, ia_len[p], ctx_l + (size_t)p * cap,
                                                        ctx_r + (size_t)p * cap, ctx_len[p], ctx_m + (size_t)c * cap,
                                                        ctx_l + (size_t)c * cap, ctx_r + (size_t)c * cap, ctx_len[c],
                                                        ia_m + (size_t)c * cap, ia_l + (size_t)c * cap, ia_r + (size_t)c * cap)) {
                        rc = -2;
                        goto out;
                    }
                    ia_len[c] = ia_len[p];
                    oracle_lexically_disambiguate(ia_m + (size_t)c * cap, ia_len[c], t->a, single + (size_t)c * cap);
                }
                done[c] = 1;
            }
        }
    rc = root;
out:
    free(done);
    free(parent);
    return rc;
}
