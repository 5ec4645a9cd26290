/* reroot_oracle.c — CPU restatement of PCG's dynamic-character RE-ROOTING pass, trees only (TEST INFRASTRUCTURE).
 *
 * Reference: lib/core/data-structures/src/Bio/Graph/PhylogeneticDAG/DynamicCharacterRerooting.hs
 *   :122-128  otherUnrootedEdges / rootEdgeReferences — the unrooted edges: the root's two children form ONE edge,
 *             every other parent-child pair is an edge;
 *   :176-199  referenceEdgeMapping — cost of rooting on edge (i,j): the root edge keeps the post-order root;
 *             otherwise localResolutionApplication (datum of i entering from j) (datum of j entering from i);
 *   :230-330  contextNodeDatum — directed subtree data, memoized: a leaf's only datum is itself; at an internal
 *             non-root node n with adjacent nodes [c1, c2, p] (children in index order, then the unrooted parent —
 *             the SIBLING when the parent is the root, :262-272) the datum entering from i is, over the rotations
 *             (i,j,k) of that list (:291-302): the post-order datum of n when i is n's original parent (:318),
 *             else localResolutionApplication (datum of j entering from n) (datum of k entering from n) (:326);
 *   :345-420  per character, the minimum over edges (ties keep every minimal edge).
 * localResolutionApplication on a dynamic character is directOptimizationPostorderPairwise
 * (DO/Internal.hs:127-140): local = DO cost, total = local + both totals; only the children's UNGAPPED MEDIANS feed
 * the alignment (DO/Pairwise/Internal.hs:364-381), so a datum here is (ungapped median string, total cost).
 * The C-linear dialect (oracle_align2d) is the aligner.
 *
 * PINNED by the reference's golden test: chel.tree re-roots to a total of 336 with local cost 49 on the minimal
 * edge (test/data-sets/golden-tests/chel-golden-test/chel_xml.golden, node 0), tests/test_oracle.py. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "do_oracle.h"

typedef struct { uint32_t *s; size_t n; int64_t total; int64_t local; int done; } datum_t;

typedef struct {
    const oracle_tcm_t *t;            /* dense C-linear aligner (NULL: the Haskell dialect below) */
    int hs_dialect, hs_kind, hs_a;    /* oracle_haskell_do parameters for large alphabets */
    const uint32_t *hs_sigma;
    uint32_t gap;
    int W;                            /* words per element (1 unless the alphabet exceeds 32 symbols) */
    uint32_t n_taxa, n_nodes, root;
    const int32_t *children;          /* [(n_taxa-1)*2], internal node x has id n_taxa + x */
    int32_t *parent;                  /* original parent, -1 for the root */
    int32_t (*adj)[3];                /* [c1, c2, unrooted parent] for internal non-root nodes */
    datum_t *down;                    /* post-order datum of every node */
    datum_t (*dir)[3];                /* dir[n][slot]: datum of n entering from adj[n][slot] */
    uint64_t cells;
} rr_t;

#define RR_MISSING ((size_t)-1)       /* length of a Missing dynamic character */

static void apply(rr_t *R, const datum_t *l, const datum_t *r, datum_t *out)
{
    const int W = R->W > 1 ? R->W : 1;
    /* handleMissingCharacter (DO/Pairwise/Internal.hs:247-259): a Missing operand costs nothing and yields the other;
     * directOptimization / algn2d (:224-233, FFI.hsc:240-248): both empty once ungapped -> Missing (toMissing), one
     * empty -> cost 0 and the medians of (element, gap). */
    const int ml = l->n == RR_MISSING, mr = r->n == RR_MISSING;
    if (ml || mr || l->n == 0 || r->n == 0) {
        out->local = 0;
        out->total = l->total + r->total;
        out->done = 1;
        if ((ml && mr) || (!ml && !mr && l->n == 0 && r->n == 0)) { out->s = NULL; out->n = RR_MISSING; return; }
        const datum_t *x = (ml || (!mr && l->n == 0)) ? r : l;          /* the operand that survives */
        out->s = malloc(sizeof(uint32_t) * (x->n + 1) * (size_t)W);
        if (ml || mr) { memcpy(out->s, x->s, sizeof(uint32_t) * x->n * (size_t)W); out->n = x->n; return; }
        out->n = 0;
        for (size_t k = 0; k < x->n; k++) {
            uint32_t md[64];
            int is_gap = 1;
            if (R->t) md[0] = R->t->median[((size_t)x->s[k] << R->t->a) + R->gap];
            else if (W == 1) oracle_overlap(R->hs_kind, R->hs_sigma, R->hs_a, x->s[k], R->gap, md);
            else {
                uint32_t gw[64] = {0};
                gw[(R->hs_a - 1) >> 5] = 1u << ((R->hs_a - 1) & 31);
                oracle_overlap_w(R->hs_kind, R->hs_sigma, R->hs_a, x->s + k * W, gw, md);
            }
            for (int w = 0; w < W; w++) {
                const uint32_t want = (W == 1) ? R->gap : ((R->hs_a - 1) >> 5) == w ? 1u << ((R->hs_a - 1) & 31) : 0u;
                if (md[w] != want) is_gap = 0;
            }
            if (!is_gap) { memcpy(out->s + out->n * W, md, sizeof(uint32_t) * (size_t)W); out->n++; }
        }
        return;
    }
    const size_t cap = (l->n + r->n + 2) * (size_t)W;
    uint32_t *g = malloc(sizeof(uint32_t) * 3 * cap);
    size_t nal = 0;
    uint64_t c = 0;
    int fo = 0;
    int64_t cost;
    if (R->t) cost = oracle_align2d(R->t, l->s, l->n, r->s, r->n, g, g + cap, g + 2 * cap, &nal, &c);
    else if (W == 1) cost = oracle_haskell_do(R->hs_dialect, R->hs_kind, R->hs_sigma, R->hs_a, l->s, l->n, r->s, r->n,
                                              g, g + cap, g + 2 * cap, &nal, &c, &fo);
    else cost = oracle_haskell_do_w(R->hs_dialect, R->hs_kind, R->hs_sigma, R->hs_a, l->s, l->n, r->s, r->n,
                                    g, g + cap, g + 2 * cap, &nal, &c, &fo);
    R->cells += c;
    out->s = malloc(sizeof(uint32_t) * (nal + 1) * (size_t)W);
    out->n = 0;
    for (size_t k = 0; k < nal; k++) {
        int is_gap = 1;                                      /* deleteGaps: the median is exactly the gap */
        for (int w = 0; w < W; w++) {
            const uint32_t want = (W == 1) ? R->gap : ((R->hs_a - 1) >> 5) == w ? 1u << ((R->hs_a - 1) & 31) : 0u;
            if (g[k * W + w] != want) is_gap = 0;
        }
        if (!is_gap) { memcpy(out->s + out->n * W, g + k * W, sizeof(uint32_t) * (size_t)W); out->n++; }
    }
    out->local = cost;
    out->total = cost + l->total + r->total;
    out->done = 1;
    free(g);
}

static const datum_t *down_of(rr_t *R, uint32_t v)
{
    datum_t *d = &R->down[v];
    if (!d->done) {
        const int32_t *ch = R->children + 2 * (size_t)(v - R->n_taxa);
        apply(R, down_of(R, (uint32_t)ch[0]), down_of(R, (uint32_t)ch[1]), d);
    }
    return d;
}

/* datum of node n entering from node `from` */
static const datum_t *entering(rr_t *R, uint32_t n, uint32_t from)
{
    if (n < R->n_taxa) return &R->down[n];                       /* leaf: HM.singleton (parentRef, n) (getCache n) */
    int slot = -1;
    for (int s = 0; s < 3; s++)
        if ((uint32_t)R->adj[n][s] == from) slot = s;
    if (slot < 0) abort();
    datum_t *d = &R->dir[n][slot];
    if (d->done) return d;
    if ((int32_t)from == R->parent[n]) {                         /* base case: the post-order datum */
        *d = *down_of(R, n);
        return d;
    }
    const uint32_t j = (uint32_t)R->adj[n][(slot + 1) % 3], k = (uint32_t)R->adj[n][(slot + 2) % 3];
    apply(R, entering(R, j, n), entering(R, k, n), d);
    return d;
}

/* One tree, one character.  leaf[v] / leaf_len[v]: ungapped element strings of the taxa.
 * Outputs (2*n_taxa - 3 edges): edge_u/edge_v endpoints, edge_cost total cost of rooting there, edge_local the
 * local cost of that rooting.  Edge 0 is the original root edge.  Returns the minimum edge cost. */
static int64_t trial_cost(rr_t *R, const datum_t *edge, const uint32_t *leaf, size_t n)
{
    /* iterativeBuild.tryEdge (app/pcg/PCG/Command/Build/Evaluate.hs:494-527): the local cost of the post-order
     * function applied to (edge datum, new leaf) */
    datum_t l = { (uint32_t *)leaf, n, 0, 0, 1 }, out;
    memset(&out, 0, sizeof out);
    apply(R, edge, &l, &out);
    free(out.s);
    return out.local;
}

int64_t oracle_reroot_trial(const oracle_tcm_t *t, uint32_t n_taxa, const int32_t *children,
                            const uint32_t *const *leaf, const size_t *leaf_len,
                            int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost, int64_t *edge_local, uint64_t *cells,
                            const uint32_t *trial, size_t trial_len, int64_t *edge_trial);

int64_t oracle_reroot(const oracle_tcm_t *t, uint32_t n_taxa, const int32_t *children,
                      const uint32_t *const *leaf, const size_t *leaf_len,
                      int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost, int64_t *edge_local, uint64_t *cells)
{
    return oracle_reroot_trial(t, n_taxa, children, leaf, leaf_len, edge_u, edge_v, edge_cost, edge_local, cells,
                               NULL, 0, NULL);
}

static int64_t reroot_run(rr_t *Rp, uint32_t n_taxa, const int32_t *children, const uint32_t *const *leaf,
                          const size_t *leaf_len, int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost,
                          int64_t *edge_local, uint64_t *cells, const uint32_t *trial, size_t trial_len,
                          int64_t *edge_trial);

/* As oracle_reroot; with `trial` != NULL also edge_trial[e] = cost increase of attaching that new leaf to edge e. */
int64_t oracle_reroot_trial(const oracle_tcm_t *t, uint32_t n_taxa, const int32_t *children,
                            const uint32_t *const *leaf, const size_t *leaf_len,
                            int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost, int64_t *edge_local, uint64_t *cells,
                            const uint32_t *trial, size_t trial_len, int64_t *edge_trial)
{
    rr_t R;
    memset(&R, 0, sizeof R);
    R.t = t; R.gap = t->gap;
    return reroot_run(&R, n_taxa, children, leaf, leaf_len, edge_u, edge_v, edge_cost, edge_local, cells, trial,
                      trial_len, edge_trial);
}

/* The same pass with every alignment done by a Haskell dialect over an alphabet of `a` symbols (dialect / metric_kind /
 * sigma as oracle_haskell_do): what PCG runs for alphabets > 8. */
int64_t oracle_reroot_haskell(int dialect, int metric_kind, const uint32_t *sigma, int a, uint32_t n_taxa,
                              const int32_t *children, const uint32_t *const *leaf, const size_t *leaf_len,
                              int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost, int64_t *edge_local)
{
    rr_t R;
    memset(&R, 0, sizeof R);
    R.t = NULL; R.hs_dialect = dialect; R.hs_kind = metric_kind; R.hs_sigma = sigma; R.hs_a = a;
    R.W = (a + 31) / 32;
    R.gap = a <= 32 ? 1u << (a - 1) : 0u;
    return reroot_run(&R, n_taxa, children, leaf, leaf_len, edge_u, edge_v, edge_cost, edge_local, NULL, NULL, 0, NULL);
}

static int64_t reroot_run(rr_t *Rp, uint32_t n_taxa, const int32_t *children, const uint32_t *const *leaf,
                          const size_t *leaf_len, int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost,
                          int64_t *edge_local, uint64_t *cells, const uint32_t *trial, size_t trial_len,
                          int64_t *edge_trial)
{
    rr_t R = *Rp;
    R.n_taxa = n_taxa; R.n_nodes = 2 * n_taxa - 1; R.children = children;
    R.parent = malloc(sizeof(int32_t) * R.n_nodes);
    R.adj = calloc(R.n_nodes, sizeof *R.adj);
    R.down = calloc(R.n_nodes, sizeof(datum_t));
    R.dir = calloc(R.n_nodes, sizeof *R.dir);
    for (uint32_t v = 0; v < R.n_nodes; v++) R.parent[v] = -1;
    for (uint32_t x = 0; x + 1 < n_taxa; x++)
        for (int s = 0; s < 2; s++) R.parent[children[2 * x + s]] = (int32_t)(n_taxa + x);
    for (uint32_t v = n_taxa; v < R.n_nodes; v++)
        if (R.parent[v] < 0) R.root = v;
    for (uint32_t v = 0; v < n_taxa; v++) {
        R.down[v].s = (uint32_t *)leaf[v]; R.down[v].n = leaf_len[v]; R.down[v].total = 0; R.down[v].done = 1;
    }
    const int32_t *rc = children + 2 * (size_t)(R.root - n_taxa);
    for (uint32_t v = n_taxa; v < R.n_nodes; v++) {
        if (v == R.root) continue;
        const int32_t *ch = children + 2 * (size_t)(v - n_taxa);
        int32_t c1 = ch[0] < ch[1] ? ch[0] : ch[1], c2 = ch[0] < ch[1] ? ch[1] : ch[0];   /* IM.keys: ascending */
        int32_t p = R.parent[v];
        if ((uint32_t)p == R.root) p = rc[0] == (int32_t)v ? rc[1] : rc[0];               /* sibling under the root */
        R.adj[v][0] = c1; R.adj[v][1] = c2; R.adj[v][2] = p;
    }
    size_t ne = 0;
    int64_t best = INT64_MAX;
    /* the root edge keeps the post-order root resolution */
    {
        const datum_t *d = down_of(&R, R.root);
        edge_u[ne] = rc[0]; edge_v[ne] = rc[1]; edge_cost[ne] = d->total; edge_local[ne] = d->local;
        if (trial) edge_trial[ne] = trial_cost(&R, d, trial, trial_len);
        ne++;
        best = d->total;
    }
    for (uint32_t i = n_taxa; i < R.n_nodes; i++) {
        if (i == R.root) continue;
        const int32_t *ch = children + 2 * (size_t)(i - n_taxa);
        for (int s = 0; s < 2; s++) {
            const uint32_t j = (uint32_t)ch[s];
            datum_t e;
            memset(&e, 0, sizeof e);
            apply(&R, entering(&R, i, j), entering(&R, j, i), &e);
            edge_u[ne] = (int32_t)i; edge_v[ne] = (int32_t)j; edge_cost[ne] = e.total; edge_local[ne] = e.local;
            if (trial) edge_trial[ne] = trial_cost(&R, &e, trial, trial_len);
            ne++;
            if (e.total < best) best = e.total;
            free(e.s);
        }
    }
    if (cells) *cells = R.cells;
    /* directed data alias post-order data in the base case: free each buffer once */
    for (uint32_t v = n_taxa; v < R.n_nodes; v++) {
        for (int s = 0; s < 3; s++)
            if (R.dir[v][s].done && R.dir[v][s].s != R.down[v].s) free(R.dir[v][s].s);
        if (R.down[v].done) free(R.down[v].s);
    }
    free(R.parent); free(R.adj); free(R.down); free(R.dir);
    return best;
}
