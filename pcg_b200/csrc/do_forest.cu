// do_forest.cu — post-order level scheduler: scores many candidate trees x dynamic characters on one GPU.
//
// Reference path being replaced (children-before-parent evaluation, one pairwise DO per internal node
// and dynamic character, tree cost = sum over characters of the root's total):
//   lib/core/data-structures/src/Bio/Graph/PhylogeneticDAG/Postorder.hs:98-158   (memoised post-order)
//   .../PhylogeneticDAG/Internal.hs:464-522                                     (generateLocalResolutions)
//   lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization/Internal.hs:109-140
//       (initializeLeaf: cost 0; directOptimizationPostorderPairwise: total = local + left + right)
//
// B200 design: nodes of equal height across ALL trees and characters form one batch ("level"); a level is
// three launches on one stream (prep: child slots -> batch descriptors; the narrow / wide DO kernels in
// ungapped-median mode; post: totals).  Everything a parent needs from a child is the child's UNGAPPED
// MEDIAN STRING (the Haskell wrapper strips gap-median triples before every alignment,
// DO/Pairwise/Internal.hs:364-381), so that is what nodes keep resident in HBM: fixed-capacity byte slots,
// leaves shared by all trees.  No host round trip between levels; one check of the overflow counters at
// the end.  Multi-GPU: ranks own character blocks of the same trees and all-reduce the per-tree sums.
//
// Re-rooting (pcgdo_forest_enable_rerooting / pcgdo_forest_score_rerooted; trees only) replaces
//   .../PhylogeneticDAG/DynamicCharacterRerooting.hs:70-420  (assignOptimalDynamicCharacterRootEdges):
// after the post-order, a PRE-ORDER sweep by depth computes for every internal non-root node n with adjacent
// nodes [c1, c2, p] (children in index order, unrooted parent = the sibling when the parent is the root) the two
// directed data "n entered from c1" = DO(n->c2 , n->p) and "n entered from c2" = DO(n->p , n->c1) (:291-326; the
// third, "entered from p", is the post-order datum), then ONE cost-only batch aligns, for every non-root edge
// (q, child), "q entered from child" with the child's post-order datum (:176-199); per character the minimum over
// the 2n-3 edges (the root edge keeps the post-order root) is the re-rooted cost (:345-420).
#include "do_ctx.h"
#include "do_metric.cuh"

#include <algorithm>
#include <cstdio>

using namespace pcgdo;

// out = slot of the result within its (tree, character) column: x for the post-order datum of internal node x,
// n_internal + 2x + s for "x entered from its s-th child (index order)", or an edge number for cost-only edge jobs;
// left / right = node refs: < n_taxa a leaf, otherwise n_taxa + slot
struct LevelItem { int32_t tree, out, left, right; };

struct pcgdo_forest {
    uint32_t n_trees = 0, n_taxa = 0, n_chars = 0, cap = 0, n_internal = 0;
    uint32_t stride = 0;                 // result slots per (tree, character) column: n_internal, or 3x with re-rooting
    bool reroot = false;
    uint32_t n_edges = 0;                // 2 * n_taxa - 3 unrooted edges per tree, edge 0 = the root edge
    std::vector<int32_t> h_children, h_edges;                 // h_edges: (u, v) per tree and edge
    std::vector<std::pair<size_t, size_t>> up_levels;         // pre-order levels (by depth) of directed-datum jobs
    std::pair<size_t, size_t> edge_items{0, 0};
    DevBuf edge_total, edge_local, char_min;                  // i64 / i32 [n_trees*n_chars*n_edges], i64 [n_trees*n_chars]
    uint32_t elem_bytes = 1;             // 1: dense-TCM alphabets (do_linear.cu); 4 * ceil(a/32): large alphabets (do_explicit.cu)
    const pcgdo_memo *memo = nullptr;    // elem_bytes 4: explicit matrix, or NULL = discrete metric over `alphabet` symbols
    int alphabet = 0, dialect = 2;
    bool keep_edges = false;             // edge data (median of rooting on each edge) kept for Wagner edge trials
    size_t trial_off = 0;                // byte offset of the trial leaf's slots in `store`
    DevBuf trial_len, trial_delta;       // u32 [n_chars], i64 [n_trees * n_edges]
    std::pair<size_t, size_t> trial_items{0, 0};
    size_t leaf_bytes = 0;               // leaf region of `store`, in ELEMENTS (as every offset below)
    DevBuf store;        // [leaf slots | internal-node slots], cap bytes each
    DevBuf leaf_len;     // u32 [n_taxa * n_chars]
    DevBuf node_len;     // u32 [n_trees * n_chars * n_internal]
    DevBuf node_local;   // i32 same
    DevBuf node_total;   // i64 same
    DevBuf items;        // LevelItem, all levels back to back
    DevBuf roots;        // i32 [n_trees] internal index of the root
    DevBuf tree_cost;    // i64 [n_trees]
    DevBuf boff1, boff2, buoff, blen1, blen2, bcost, bulen;   // per-level batch descriptors
    std::vector<std::pair<size_t, size_t>> levels;            // (first item, count)
    size_t max_level_items = 0;          // items per launch (fixed at creation: room for the widest possible level)
    std::vector<int32_t> h_roots;
    // candidate-tree turnover (pcgdo_forest_set_trees): topologies are levelled ON THE DEVICE into the existing buffers
    DevBuf d_children;   // i32 [n_trees][n_internal][2]
    DevBuf d_height;     // i32 [n_trees][n_internal]
    DevBuf d_hist;       // u32 [n_taxa + 2 | n_taxa + 2 | 4]: nodes per height, scatter cursors, error flags
    size_t items_cap = 0;                // LevelItem entries allocated
    bool keep_requested = false;
};

namespace {

struct ForestDev {
    uint32_t n_taxa, n_chars, n_internal, cap;        // n_internal = result slots per column (stride)
    uint64_t leaf_bytes;
    uint64_t trial_off;                               // refs >= n_taxa + n_internal: the trial leaf
    const uint32_t *trial_len;
    const uint32_t *leaf_len;
    uint32_t *node_len;
    int32_t *node_local;
    int64_t *node_total;
    uint32_t *store32;                                // the node store as 32-bit words (slots are 4-element aligned)
    uint32_t elem_bytes;
};

constexpr uint32_t kMissing = PCGDO_MISSING;          // length of a Missing dynamic character

__device__ __forceinline__ uint64_t slot_of(const ForestDev &f, uint32_t tree, uint32_t c, int32_t node)
{
    if ((uint32_t)node < f.n_taxa) return ((uint64_t)node * f.n_chars + c) * f.cap;
    if ((uint32_t)node >= f.n_taxa + f.n_internal) return f.trial_off + (uint64_t)c * f.cap;
    return f.leaf_bytes + (((uint64_t)tree * f.n_chars + c) * f.n_internal + (uint32_t)(node - f.n_taxa)) * f.cap;
}
__device__ __forceinline__ uint64_t nidx(const ForestDev &f, uint32_t tree, uint32_t c, uint32_t x)
{
    return ((uint64_t)tree * f.n_chars + c) * f.n_internal + x;
}

__device__ __forceinline__ void child_lengths(const ForestDev &f, const LevelItem &it, uint32_t c, uint32_t &l1, uint32_t &l2)
{
    l1 = (uint32_t)it.left < f.n_taxa ? f.leaf_len[(uint64_t)it.left * f.n_chars + c]
                                      : f.node_len[nidx(f, it.tree, c, it.left - f.n_taxa)];
    l2 = (uint32_t)it.right < f.n_taxa ? f.leaf_len[(uint64_t)it.right * f.n_chars + c]
         : (uint32_t)it.right >= f.n_taxa + f.n_internal ? f.trial_len[c]
                                                         : f.node_len[nidx(f, it.tree, c, it.right - f.n_taxa)];
}

__global__ void forest_prep_kernel(ForestDev f, const LevelItem *items, uint32_t n_items, uint64_t *off1,
                                   uint64_t *off2, uint64_t *uoff, uint32_t *len1, uint32_t *len2)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)n_items * f.n_chars) return;
    const LevelItem it = items[k / f.n_chars];
    const uint32_t c = (uint32_t)(k % f.n_chars);
    off1[k] = slot_of(f, it.tree, c, it.left);
    off2[k] = slot_of(f, it.tree, c, it.right);
    uoff[k] = slot_of(f, it.tree, c, (int32_t)(f.n_taxa + it.out));
    uint32_t l1, l2;
    child_lengths(f, it, c, l1, l2);
    if (l1 == kMissing || l2 == kMissing) l1 = l2 = 0;      // nothing to align: forest_special_kernel supplies the result
    len1[k] = l1;
    len2[k] = l2;
}

// The cases the reference settles before (or instead of) aligning — `handleMissingCharacter`
// (DO/Pairwise/Internal.hs:247-259) and the empty-after-ungapping branches of `directOptimization` / `algn2d`
// (:224-233, DO/Pairwise/FFI.hsc:240-248) — applied to the batch after the aligner ran and before the totals are taken:
//   a Missing child            -> cost 0, the node is the other child's string (Missing if both are);
//   both children empty        -> cost 0, the node is Missing (`toMissing`);
//   one child empty            -> cost 0 (the aligner's medians of (x, gap) stand; the reference charges nothing).
// data = the batch keeps node strings (post-order / directed data / kept edge data), else it is cost-only.
__global__ void forest_special_kernel(ForestDev f, const LevelItem *items, uint32_t n_items, const uint64_t *off1,
                                      const uint64_t *off2, const uint64_t *uoff, int32_t *cost, uint32_t *ulen, int data)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)n_items * f.n_chars) return;
    const LevelItem it = items[k / f.n_chars];
    const uint32_t c = (uint32_t)(k % f.n_chars);
    uint32_t l1, l2;
    child_lengths(f, it, c, l1, l2);
    const bool m1 = l1 == kMissing, m2 = l2 == kMissing;
    if (!m1 && !m2 && l1 != 0 && l2 != 0) return;
    cost[k] = 0;
    if (!data) return;
    if (m1 && m2) { ulen[k] = kMissing; return; }
    if (m1 || m2) {
        const uint32_t n = m1 ? l2 : l1;
        const uint32_t *src = f.store32 + (m1 ? off2[k] : off1[k]) * f.elem_bytes / 4;
        uint32_t *dst = f.store32 + uoff[k] * f.elem_bytes / 4;
        const uint32_t words = (uint32_t)(((uint64_t)n * f.elem_bytes + 3) / 4);
        for (uint32_t w = 0; w < words; ++w) dst[w] = src[w];
        ulen[k] = n;
        return;
    }
    if (l1 == 0 && l2 == 0) ulen[k] = kMissing;
}

__global__ void forest_post_kernel(ForestDev f, const LevelItem *items, uint32_t n_items, const int32_t *cost,
                                   const uint32_t *ulen)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)n_items * f.n_chars) return;
    const LevelItem it = items[k / f.n_chars];
    const uint32_t c = (uint32_t)(k % f.n_chars);
    const uint64_t me = nidx(f, it.tree, c, it.out);
    int64_t tot = cost[k];
    if ((uint32_t)it.left >= f.n_taxa) tot += f.node_total[nidx(f, it.tree, c, it.left - f.n_taxa)];
    if ((uint32_t)it.right >= f.n_taxa) tot += f.node_total[nidx(f, it.tree, c, it.right - f.n_taxa)];
    f.node_local[me] = cost[k];
    f.node_total[me] = tot;
    f.node_len[me] = ulen[k];
}

// tree cost = sum over this rank's characters of the root's total (unit weights; weighting is the caller's)
__global__ void forest_reduce_kernel(ForestDev f, const int32_t *roots, uint32_t n_trees, int64_t *tree_cost)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    int64_t s = 0;
    for (uint32_t c = 0; c < f.n_chars; ++c) s += f.node_total[nidx(f, t, c, (uint32_t)roots[t])];
    tree_cost[t] = s;
}

// cost-only edge jobs: total of rooting on that edge = local + both data's totals
__global__ void forest_edge_post_kernel(ForestDev f, const LevelItem *items, uint32_t n_items, const int32_t *cost,
                                        uint32_t n_edges, int64_t *edge_total, int32_t *edge_local)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)n_items * f.n_chars) return;
    const LevelItem it = items[k / f.n_chars];
    const uint32_t c = (uint32_t)(k % f.n_chars);
    int64_t tot = cost[k];
    if ((uint32_t)it.left >= f.n_taxa) tot += f.node_total[nidx(f, it.tree, c, it.left - f.n_taxa)];
    if ((uint32_t)it.right >= f.n_taxa) tot += f.node_total[nidx(f, it.tree, c, it.right - f.n_taxa)];
    const uint64_t e = ((uint64_t)it.tree * f.n_chars + c) * n_edges + (uint32_t)it.out;
    edge_total[e] = tot;
    edge_local[e] = cost[k];
}

// edge jobs that keep their datum: totals per edge as above, plus the datum's length in its slot
__global__ void forest_edge_keep_kernel(ForestDev f, const LevelItem *items, uint32_t n_items, const int32_t *cost,
                                        const uint32_t *ulen, uint32_t n_edges, uint32_t edge_slot0, int64_t *edge_total,
                                        int32_t *edge_local)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)n_items * f.n_chars) return;
    const LevelItem it = items[k / f.n_chars];
    const uint32_t c = (uint32_t)(k % f.n_chars);
    int64_t tot = cost[k];
    if ((uint32_t)it.left >= f.n_taxa) tot += f.node_total[nidx(f, it.tree, c, it.left - f.n_taxa)];
    if ((uint32_t)it.right >= f.n_taxa) tot += f.node_total[nidx(f, it.tree, c, it.right - f.n_taxa)];
    const uint32_t edge = (uint32_t)it.out - edge_slot0;
    const uint64_t e = ((uint64_t)it.tree * f.n_chars + c) * n_edges + edge;
    edge_total[e] = tot;
    edge_local[e] = cost[k];
    f.node_len[nidx(f, it.tree, c, it.out)] = ulen[k];
}

// Wagner edge trial: delta[tree][edge] = sum over characters of the local cost of (edge datum, new leaf)
__global__ void forest_trial_sum_kernel(const LevelItem *items, uint32_t n_items, uint32_t n_chars, uint32_t n_edges,
                                        const int32_t *cost, int64_t *delta)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_items) return;
    const LevelItem it = items[k];
    int64_t s = 0;
    for (uint32_t c = 0; c < n_chars; ++c) s += cost[(uint64_t)k * n_chars + c];
    delta[(uint64_t)it.tree * n_edges + (uint32_t)it.out] = s;
}

// per (tree, character): the root edge keeps the post-order root; minimum over all edges
__global__ void forest_edge_min_kernel(ForestDev f, const int32_t *roots, uint32_t n_trees, uint32_t n_edges,
                                       int64_t *edge_total, int32_t *edge_local, int64_t *char_min)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)n_trees * f.n_chars) return;
    const uint32_t t = (uint32_t)(k / f.n_chars), c = (uint32_t)(k % f.n_chars);
    const uint64_t r = nidx(f, t, c, (uint32_t)roots[t]);
    int64_t *et = edge_total + k * n_edges;
    et[0] = f.node_total[r];
    edge_local[k * n_edges] = f.node_local[r];
    int64_t best = et[0];
    for (uint32_t e = 1; e < n_edges; ++e) best = min(best, et[e]);
    char_min[k] = best;
}

__global__ void forest_reduce_min_kernel(uint32_t n_chars, const int64_t *char_min, uint32_t n_trees, int64_t *tree_cost)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    int64_t s = 0;
    for (uint32_t c = 0; c < n_chars; ++c) s += char_min[(uint64_t)t * n_chars + c];
    tree_cost[t] = s;
}

// ---- levelling of a fresh set of topologies (pcgdo_forest_set_trees).  One CTA per tree: heights by relaxation in
// shared memory (a node is settled once both children are: at most `depth` rounds), validation (every id in range and
// the child of exactly one node, one root, no cycle), the root, and this tree's share of the per-height histogram.
__global__ void forest_height_kernel(const int32_t *children, uint32_t n_taxa, int32_t *height_out, int32_t *roots,
                                     unsigned int *hist, unsigned int *err)
{
    extern __shared__ int32_t fh_smem[];
    const uint32_t ni = n_taxa - 1, nn = 2 * n_taxa - 1, t = blockIdx.x;
    int32_t *height = fh_smem;                 // [nn]
    int32_t *uses = fh_smem + nn;              // [nn] times a node appears as a child
    int32_t *lh = uses + nn;                   // [n_taxa] this tree's histogram
    __shared__ int s_changed, s_bad, s_root, s_nroot;
    const int32_t *ch = children + (size_t)t * ni * 2;
    for (uint32_t v = threadIdx.x; v < nn; v += blockDim.x) { height[v] = v < n_taxa ? 0 : -1; uses[v] = 0; }
    for (uint32_t v = threadIdx.x; v < n_taxa; v += blockDim.x) lh[v] = 0;
    if (threadIdx.x == 0) { s_bad = 0; s_root = -1; s_nroot = 0; }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < 2 * ni; e += blockDim.x) {
        const int32_t c = ch[e];
        if (c < 0 || (uint32_t)c >= nn) s_bad = 1;
        else if (atomicAdd(&uses[c], 1) != 0) s_bad = 1;
    }
    __syncthreads();
    if (!s_bad) {
        for (uint32_t round = 0; round <= ni; ++round) {
            if (threadIdx.x == 0) s_changed = 0;
            __syncthreads();
            for (uint32_t x = threadIdx.x; x < ni; x += blockDim.x) {
                if (height[n_taxa + x] >= 0) continue;
                const int32_t hl = height[ch[2 * x]], hr = height[ch[2 * x + 1]];
                if (hl >= 0 && hr >= 0) { height[n_taxa + x] = 1 + max(hl, hr); s_changed = 1; }
            }
            __syncthreads();
            const int changed = s_changed;
            __syncthreads();
            if (!changed) break;
        }
        for (uint32_t x = threadIdx.x; x < ni; x += blockDim.x) {
            const int32_t h = height[n_taxa + x];
            if (h < 0) s_bad = 1;                                     // a cycle
            else atomicAdd(&lh[h], 1);
            if (!uses[n_taxa + x]) { atomicAdd(&s_nroot, 1); s_root = (int)x; }
            height_out[(size_t)t * ni + x] = h;
        }
    }
    __syncthreads();
    if (s_bad || s_nroot != 1) { if (threadIdx.x == 0) atomicAdd(err, 1u); return; }
    if (threadIdx.x == 0) roots[t] = s_root;
    for (uint32_t h = threadIdx.x; h < n_taxa; h += blockDim.x)
        if (lh[h]) atomicAdd(&hist[h], (unsigned int)lh[h]);
}

// Items grouped by height: item position = start of its height's range + a ticket from that height's cursor.
__global__ void forest_scatter_kernel(const int32_t *children, const int32_t *height, uint32_t n_trees, uint32_t n_taxa,
                                      unsigned int *cursor, LevelItem *items)
{
    const uint32_t ni = n_taxa - 1;
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)n_trees * ni) return;
    const int32_t h = height[k];
    const unsigned int pos = atomicAdd(&cursor[h], 1u);
    const uint32_t t = (uint32_t)(k / ni), x = (uint32_t)(k % ni);
    items[pos] = {(int32_t)t, (int32_t)x, children[2 * k], children[2 * k + 1]};
}

// DP cells of the reference's band over the alignments of the last post-order sweep (roofline accounting): one thread
// per (item, character), lengths as resident after the sweep.
__global__ void forest_cells_kernel(ForestDev f, const LevelItem *items, uint64_t n_items, unsigned long long *total)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    if (k < n_items * f.n_chars) {
        const LevelItem it = items[k / f.n_chars];
        uint32_t l1, l2;
        child_lengths(f, it, (uint32_t)(k % f.n_chars), l1, l2);
        if (l1 != kMissing && l2 != kMissing && l1 && l2) {
            const Geom g = make_geom((int)max(l1, l2) + 1, (int)min(l1, l2) + 1);
            c = band_cells(g, 0, 1);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}

ForestDev dev_view(const pcgdo_forest *F)
{
    ForestDev f;
    f.n_taxa = F->n_taxa; f.n_chars = F->n_chars; f.n_internal = F->stride; f.cap = F->cap;
    f.leaf_bytes = F->leaf_bytes;
    f.trial_off = F->trial_off;
    f.trial_len = (const uint32_t *)F->trial_len.p;
    f.leaf_len = (const uint32_t *)F->leaf_len.p;
    f.node_len = (uint32_t *)F->node_len.p;
    f.node_local = (int32_t *)F->node_local.p;
    f.node_total = (int64_t *)F->node_total.p;
    f.store32 = (uint32_t *)F->store.p;
    f.elem_bytes = F->elem_bytes;
    return f;
}

}  // namespace

static int forest_create_impl(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars, uint32_t node_cap,
                              const int32_t *children, uint32_t elem_bytes, pcgdo_forest **out);

extern "C" int pcgdo_forest_create(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars,
                                   uint32_t node_cap, const int32_t *children, pcgdo_forest **out)
{
    return forest_create_impl(ctx, n_trees, n_taxa, n_chars, node_cap, children, 1, out);
}

// Large alphabets: 32-bit elements, aligned by the Haskell dialects of do_explicit.cu.  memo = NULL selects the
// discrete metric over `alphabet` symbols; dialect as pcgdo_metric_batch_* (2 = unboxedUkkonenFullSpaceDO).
extern "C" int pcgdo_forest_create32(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars,
                                     uint32_t node_cap, const int32_t *children, const pcgdo_memo *memo, int alphabet,
                                     int dialect, pcgdo_forest **out)
{
    if (dialect < 0 || dialect > 2 || (!memo && (alphabet < 2 || alphabet > PCGDO_MAX_ALPHABET))) return PCGDO_ERR_ARG;
    // elements are ceil(a / 32) 32-bit words (one word up to 32 symbols)
    const uint32_t a = memo ? pcgdo_memo_alphabet(memo) : (uint32_t)alphabet;
    int rc = forest_create_impl(ctx, n_trees, n_taxa, n_chars, node_cap, children, 4u * ((a + 31u) / 32u), out);
    if (rc) return rc;
    (*out)->memo = memo;
    (*out)->alphabet = alphabet;
    (*out)->dialect = dialect;
    return PCGDO_OK;
}

static int forest_build_reroot(pcgdo_ctx *ctx, pcgdo_forest *F);

// Levels a set of topologies into F's item array: heights, validation and grouping by height run on the device
// (forest_height_kernel / forest_scatter_kernel); the host reads back one small histogram to know the level sizes.
static int forest_level_trees(pcgdo_ctx *ctx, pcgdo_forest *F, const int32_t *children)
{
    const uint32_t n_trees = F->n_trees, n_taxa = F->n_taxa, ni = F->n_internal;
    const size_t n_ch = (size_t)n_trees * ni * 2;
    if (children != F->h_children.data()) F->h_children.assign(children, children + n_ch);
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(F->d_children.p, F->h_children.data(), n_ch * 4, cudaMemcpyHostToDevice, st));
    unsigned int *hist = (unsigned int *)F->d_hist.p, *cursor = hist + (n_taxa + 2), *err = cursor + (n_taxa + 2);
    CK(cudaMemsetAsync(hist, 0, (2 * (size_t)(n_taxa + 2) + 4) * 4, st));
    const size_t smem = ((size_t)2 * (2 * n_taxa - 1) + n_taxa) * 4;
    if (smem > 48 * 1024) {
        static DeviceOnce once;
        int dev = 0;
        if (once.pending(&dev)) {
            CK(cudaFuncSetAttribute(forest_height_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            once.mark(dev);
        }
    }
    forest_height_kernel<<<n_trees, 256, smem, st>>>((const int32_t *)F->d_children.p, n_taxa, (int32_t *)F->d_height.p,
                                                     (int32_t *)F->roots.p, hist, err);
    CK(cudaGetLastError());
    std::vector<unsigned int> h_hist(n_taxa + 2);
    unsigned int h_err = 0;
    F->h_roots.resize(n_trees);
    CK(cudaMemcpyAsync(h_hist.data(), hist, (size_t)(n_taxa + 2) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(F->h_roots.data(), F->roots.p, (size_t)n_trees * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_err) {
        ctx->err = "forest: children[] is not a set of binary trees over the taxa (bad id, shared child, cycle, or several roots)";
        return PCGDO_ERR_ARG;
    }
    F->levels.clear();
    std::vector<unsigned int> h_cursor(n_taxa + 2, 0u);
    size_t at = 0;
    for (uint32_t h = 1; h < n_taxa; ++h) {
        h_cursor[h] = (unsigned int)at;
        if (!h_hist[h]) continue;
        F->levels.push_back({at, (size_t)h_hist[h]});
        at += h_hist[h];
    }
    if (at != (size_t)n_trees * ni) {
        ctx->err = "forest: internal error levelling the trees";
        return PCGDO_ERR_ARG;
    }
    CK(cudaMemcpyAsync(cursor, h_cursor.data(), (size_t)(n_taxa + 2) * 4, cudaMemcpyHostToDevice, st));
    const uint64_t n_nodes = (uint64_t)n_trees * ni;
    forest_scatter_kernel<<<(uint32_t)((n_nodes + 255) / 256), 256, 0, st>>>(
        (const int32_t *)F->d_children.p, (const int32_t *)F->d_height.p, n_trees, n_taxa, cursor, (LevelItem *)F->items.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return PCGDO_OK;
}

static int forest_create_impl(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars, uint32_t node_cap,
                              const int32_t *children, uint32_t elem_bytes, pcgdo_forest **out)
{
    if (!ctx || !children || !out || n_trees == 0 || n_taxa < 2 || n_chars == 0 || node_cap < 4) return PCGDO_ERR_ARG;
    if (((size_t)2 * (2 * n_taxa - 1) + n_taxa) * 4 > 227 * 1024) {
        ctx->err = "forest: more taxa than the levelling kernel's shared memory holds (about 11 000)";
        return PCGDO_ERR_UNSUPPORTED;
    }
    CK(cudaSetDevice(ctx->device));
    pcgdo_forest *F = new pcgdo_forest();
    F->elem_bytes = elem_bytes;
    F->n_trees = n_trees; F->n_taxa = n_taxa; F->n_chars = n_chars;
    F->cap = (node_cap + 3u) & ~3u;
    F->n_internal = n_taxa - 1;
    F->stride = F->n_internal;
    F->n_edges = n_taxa >= 3 ? 2 * n_taxa - 3 : 1;
    const uint32_t ni = F->n_internal;
    // Everything is sized for ANY set of n_trees topologies over these taxa, so that pcgdo_forest_set_trees never
    // allocates: a launch holds at most n_trees * ceil(n_taxa / 2) items (no level of a binary tree is wider; wider
    // batches — the edge jobs of the re-rooting pass — are cut into launches of that size).
    F->max_level_items = (size_t)n_trees * ((n_taxa + 1) / 2);
    F->leaf_bytes = (size_t)n_taxa * n_chars * F->cap;
    const size_t n_slots = (size_t)n_trees * n_chars * ni;
    F->items_cap = (size_t)n_trees * ni;
    int rc;
    auto fail = [&](int r) { pcgdo_forest_destroy(ctx, F); return r; };
    if ((rc = pcgdo_ensure(ctx, F->store, (F->leaf_bytes + n_slots * F->cap + 64) * F->elem_bytes))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->leaf_len, (size_t)n_taxa * n_chars * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->node_len, n_slots * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->node_local, n_slots * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->node_total, n_slots * 8))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->items, F->items_cap * sizeof(LevelItem)))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->roots, (size_t)n_trees * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->tree_cost, (size_t)n_trees * 8))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->d_children, (size_t)n_trees * ni * 2 * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->d_height, (size_t)n_trees * ni * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->d_hist, (2 * (size_t)(n_taxa + 2) + 4) * 4 + 16))) return fail(rc);
    const size_t mb = F->max_level_items * n_chars;
    if ((rc = pcgdo_ensure(ctx, F->boff1, mb * 8)) || (rc = pcgdo_ensure(ctx, F->boff2, mb * 8)) ||
        (rc = pcgdo_ensure(ctx, F->buoff, mb * 8)) || (rc = pcgdo_ensure(ctx, F->blen1, mb * 4)) ||
        (rc = pcgdo_ensure(ctx, F->blen2, mb * 4)) || (rc = pcgdo_ensure(ctx, F->bcost, mb * 4)) ||
        (rc = pcgdo_ensure(ctx, F->bulen, mb * 4)))
        return fail(rc);
    if ((rc = forest_level_trees(ctx, F, children))) return fail(rc);
    *out = F;
    return PCGDO_OK;
}

// Candidate-tree turnover: a fresh set of n_trees topologies over the same taxa (same layout as at creation) replaces
// the current one.  Nothing is allocated and the leaves stay resident; the next sweep scores the new trees.  This is
// what a tree search does between sweeps (the Wagner build and the branch swappers of
// app/pcg/PCG/Command/Build/Evaluate.hs:494-527 produce new candidate topologies over fixed leaves).
extern "C" int pcgdo_forest_set_trees(pcgdo_ctx *ctx, pcgdo_forest *F, const int32_t *children)
{
    if (!ctx || !F || !children) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = forest_level_trees(ctx, F, children))) return rc;
    if (F->reroot) return forest_build_reroot(ctx, F);
    return PCGDO_OK;
}

// DP cells the reference's band fills over the n_trees * n_chars * (n_taxa - 1) alignments of the last post-order
// sweep (lengths as resident after it): the numerator of the sweep's GCUPS / roofline figures.
extern "C" int pcgdo_forest_cells(pcgdo_ctx *ctx, pcgdo_forest *F, uint64_t *cells_out)
{
    if (!ctx || !F || !cells_out) return PCGDO_ERR_ARG;
    if (F->elem_bytes != 1) {
        ctx->err = "pcgdo_forest_cells: defined for dense-TCM (8-bit) forests, whose band is a function of the lengths";
        return PCGDO_ERR_UNSUPPORTED;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    unsigned long long *d_tot = (unsigned long long *)((unsigned int *)F->d_hist.p + 2 * (size_t)(F->n_taxa + 2) + 4);
    d_tot = (unsigned long long *)(((uintptr_t)d_tot + 7) & ~(uintptr_t)7);
    CK(cudaMemsetAsync(d_tot, 0, 8, st));
    const ForestDev f = dev_view(F);
    const uint64_t n_items = (uint64_t)F->n_trees * F->n_internal, n = n_items * F->n_chars;
    forest_cells_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(f, (const LevelItem *)F->items.p, n_items, d_tot);
    CK(cudaGetLastError());
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, d_tot, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *cells_out = h;
    return PCGDO_OK;
}

extern "C" void pcgdo_forest_destroy(pcgdo_ctx *ctx, pcgdo_forest *F)
{
    if (!F) return;
    if (ctx) cudaSetDevice(ctx->device);
    DevBuf *bufs[] = {&F->store, &F->leaf_len, &F->node_len, &F->node_local, &F->node_total, &F->items, &F->roots,
                      &F->tree_cost, &F->boff1, &F->boff2, &F->buoff, &F->blen1, &F->blen2, &F->bcost, &F->bulen,
                      &F->edge_total, &F->edge_local, &F->char_min, &F->trial_len, &F->trial_delta, &F->d_children,
                      &F->d_height, &F->d_hist};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    delete F;
}

// Leaf strings (host): character c of taxon v = elems[off[v*n_chars+c] .. +len[...]) — one byte per element,
// already free of gap elements (leaves are ungapped sequences).
static int forest_set_leaves_impl(pcgdo_ctx *ctx, pcgdo_forest *F, const void *elems, const uint64_t *off, const uint32_t *len,
                                  uint32_t elem_bytes)
{
    if (!ctx || !F || !elems || !off || !len) return PCGDO_ERR_ARG;
    if (F->elem_bytes != elem_bytes) {
        ctx->err = "forest: element width of the leaves does not match the forest";
        return PCGDO_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)F->n_taxa * F->n_chars, eb = elem_bytes;
    std::vector<uint8_t> stage(n * F->cap * eb, 0);
    for (size_t i = 0; i < n; ++i) {
        if (len[i] == PCGDO_MISSING) continue;           // a Missing character: nothing to store
        if (len[i] > F->cap) {
            ctx->err = "forest: a leaf string is longer than the node capacity";
            return PCGDO_ERR_ARG;
        }
        const uint8_t *src = (const uint8_t *)elems + off[i] * eb;
        std::copy(src, src + (size_t)len[i] * eb, stage.begin() + i * F->cap * eb);
    }
    CK(cudaMemcpy(F->store.p, stage.data(), stage.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(F->leaf_len.p, len, n * 4, cudaMemcpyHostToDevice));
    return PCGDO_OK;
}

extern "C" int pcgdo_forest_set_leaves(pcgdo_ctx *ctx, pcgdo_forest *F, const uint8_t *elems, const uint64_t *off,
                                       const uint32_t *len)
{
    return forest_set_leaves_impl(ctx, F, elems, off, len, 1);
}

extern "C" int pcgdo_forest_set_leaves32(pcgdo_ctx *ctx, pcgdo_forest *F, const uint32_t *elems, const uint64_t *off,
                                         const uint32_t *len)
{
    return forest_set_leaves_impl(ctx, F, elems, off, len, F ? F->elem_bytes : 4);
}

// Enqueue one batch of forest jobs: mode 0 = data (ungapped median of the result kept in its slot, totals
// accumulated), mode 1 = cost-only edge jobs.
static int forest_run_items(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, pcgdo_forest *F, cudaStream_t st, size_t first_item,
                            size_t count, int mode, bool &first, LinearParams &p)
{
    const ForestDev f = dev_view(F);
    const size_t chunk = F->max_level_items ? F->max_level_items : count;
    for (size_t c0 = 0; c0 < count; c0 += chunk) {
        const LevelItem *items = (const LevelItem *)F->items.p + first_item + c0;
        const uint32_t n_items = (uint32_t)std::min(chunk, count - c0);
        const uint64_t n_pairs = (uint64_t)n_items * F->n_chars;
        const uint32_t blocks = (uint32_t)((n_pairs + 255) / 256);
        forest_prep_kernel<<<blocks, 256, 0, st>>>(f, items, n_items, (uint64_t *)F->boff1.p, (uint64_t *)F->boff2.p,
                                                   (uint64_t *)F->buoff.p, (uint32_t *)F->blen1.p,
                                                   (uint32_t *)F->blen2.p);
        CK(cudaGetLastError());
        int rc;
        if (F->elem_bytes >= 4) {
            MetricBatchDev md{};
            md.n_pairs = (uint32_t)n_pairs;
            md.elems = (const uint32_t *)F->store.p;
            md.off1 = (const uint64_t *)F->boff1.p; md.off2 = (const uint64_t *)F->boff2.p;
            md.len1 = (const uint32_t *)F->blen1.p; md.len2 = (const uint32_t *)F->blen2.p;
            md.cost = (int32_t *)F->bcost.p;
            if (mode == 0 || mode == 2) {
                md.out_ungapped = (uint32_t *)F->store.p;
                md.out_uoff = (const uint64_t *)F->buoff.p;
                md.out_ulen = (uint32_t *)F->bulen.p;
                md.ucap = F->cap;
            }
            unsigned int *cnt = nullptr;
            rc = pcgdo_enqueue_table(ctx, F->memo, F->alphabet, F->dialect, md, st, first, &cnt);
            p.counter = cnt;
        } else {
            BatchDev bd{};
            bd.n_pairs = (uint32_t)n_pairs;
            bd.elems = (const uint8_t *)F->store.p;
            bd.off1 = (const uint64_t *)F->boff1.p; bd.off2 = (const uint64_t *)F->boff2.p;
            bd.len1 = (const uint32_t *)F->blen1.p; bd.len2 = (const uint32_t *)F->blen2.p;
            bd.cost = (int32_t *)F->bcost.p;
            if (mode == 0 || mode == 2) {
                bd.out_ungapped = (uint8_t *)F->store.p;
                bd.out_uoff = (const uint64_t *)F->buoff.p;
                bd.out_ulen = (uint32_t *)F->bulen.p;
                bd.ucap = F->cap;
            }
            if (!tcm) {
                ctx->err = "forest: a dense TCM is required for 8-bit forests";
                return PCGDO_ERR_ARG;
            }
            rc = pcgdo_enqueue_linear(ctx, tcm, bd, st, first, &p, nullptr);
        }
        if (rc) return rc;
        first = false;
        forest_special_kernel<<<blocks, 256, 0, st>>>(f, items, n_items, (const uint64_t *)F->boff1.p, (const uint64_t *)F->boff2.p,
                                                      (const uint64_t *)F->buoff.p, (int32_t *)F->bcost.p,
                                                      (uint32_t *)F->bulen.p, mode == 0 || mode == 2);
        CK(cudaGetLastError());
        if (mode == 0)
            forest_post_kernel<<<blocks, 256, 0, st>>>(f, items, n_items, (const int32_t *)F->bcost.p,
                                                       (const uint32_t *)F->bulen.p);
        else if (mode == 1)
            forest_edge_post_kernel<<<blocks, 256, 0, st>>>(f, items, n_items, (const int32_t *)F->bcost.p, F->n_edges,
                                                            (int64_t *)F->edge_total.p, (int32_t *)F->edge_local.p);
        else if (mode == 2)
            forest_edge_keep_kernel<<<blocks, 256, 0, st>>>(f, items, n_items, (const int32_t *)F->bcost.p,
                                                            (const uint32_t *)F->bulen.p, F->n_edges, 3 * F->n_internal,
                                                            (int64_t *)F->edge_total.p, (int32_t *)F->edge_local.p);
        else
            forest_trial_sum_kernel<<<(n_items + 255) / 256, 256, 0, st>>>(items, n_items, F->n_chars, F->n_edges,
                                                                           (const int32_t *)F->bcost.p,
                                                                           (int64_t *)F->trial_delta.p);
        CK(cudaGetLastError());
        ctx->last_launches += 3;
    }
    return PCGDO_OK;
}

static int forest_check(pcgdo_ctx *ctx, pcgdo_forest *F, const LinearParams &p, cudaStream_t st)
{
    // one look at the failure counters for the whole sweep: pairs beyond the warp kernels' limits, wide pairs
    // with the wide kernel disabled, node strings beyond their slot
    unsigned int h[11] = {0};
    CK(cudaMemcpyAsync(h, p.counter, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (F->elem_bytes >= 4) {           // table kernel: [8] pairs beyond its limits, [10] node strings beyond their slot
        h[9] = h[10];
        h[3] = 0;
    }
    if (h[8] || h[9] || (ctx->max_b <= 8 && h[3])) {
        char msg[200];
        snprintf(msg, sizeof msg,
                 "forest sweep: %u pair(s) beyond the warp kernels' limits, %u node string(s) beyond node_cap=%u "
                 "(raise max_len / max_b / node_cap)", h[8] + (ctx->max_b <= 8 ? h[3] : 0u), h[9], F->cap);
        ctx->err = msg;
        return PCGDO_ERR_OVERFLOW;
    }
    return PCGDO_OK;
}

// One post-order sweep over all trees and this forest's characters; per-tree cost = sum of the root totals.
extern "C" int pcgdo_forest_score(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, pcgdo_forest *F, void *cuda_stream)
{
    if (!ctx || !F || (!tcm && F->elem_bytes == 1)) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const ForestDev f = dev_view(F);
    ctx->last_launches = 0;
    CK(cudaEventRecord(ctx->ev0, st));
    bool first = true;
    LinearParams p{};
    int rc;
    for (const auto &lv : F->levels)
        if ((rc = forest_run_items(ctx, tcm, F, st, lv.first, lv.second, 0, first, p))) return rc;
    forest_reduce_kernel<<<(F->n_trees + 127) / 128, 128, 0, st>>>(f, (const int32_t *)F->roots.p, F->n_trees,
                                                                  (int64_t *)F->tree_cost.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    return forest_check(ctx, F, p, st);
}

// Builds the pre-order (directed data) and edge job lists of the CURRENT topologies behind the post-order items.
// Their sizes do not depend on the topologies (2 directed data per internal non-root node, one job per non-root
// edge, one trial job per edge), so the item array allocated by pcgdo_forest_enable_rerooting always fits.
static int forest_build_reroot(pcgdo_ctx *ctx, pcgdo_forest *F)
{
    const uint32_t n_taxa = F->n_taxa, ni = F->n_internal, n_nodes = 2 * n_taxa - 1;
    std::vector<std::vector<LevelItem>> by_depth;
    std::vector<LevelItem> edges;
    F->h_edges.assign((size_t)F->n_trees * F->n_edges * 2, 0);
    std::vector<int32_t> parent(n_nodes), side(n_nodes), depth(n_nodes), order;
    for (uint32_t t = 0; t < F->n_trees; ++t) {
        const int32_t *ch = F->h_children.data() + (size_t)t * ni * 2;
        const int32_t root = (int32_t)(n_taxa + (uint32_t)F->h_roots[t]);
        std::fill(parent.begin(), parent.end(), -1);
        for (uint32_t x = 0; x < ni; ++x) {
            // "s-th child in index order" (IM.keys of childRefs is ascending)
            const int32_t c0 = std::min(ch[2 * x], ch[2 * x + 1]), c1 = std::max(ch[2 * x], ch[2 * x + 1]);
            parent[c0] = (int32_t)(n_taxa + x); side[c0] = 0;
            parent[c1] = (int32_t)(n_taxa + x); side[c1] = 1;
        }
        // depths, top-down
        order.clear();
        order.push_back(root);
        depth[root] = 0;
        for (size_t q = 0; q < order.size(); ++q) {
            const int32_t v = order[q];
            if ((uint32_t)v < n_taxa) continue;
            for (int s = 0; s < 2; ++s) {
                const int32_t c = ch[2 * (v - n_taxa) + s];
                depth[c] = depth[v] + 1;
                order.push_back(c);
            }
        }
        auto up_ref = [&](int32_t q, int s) { return (int32_t)(n_taxa + ni + 2 * (uint32_t)(q - n_taxa) + (uint32_t)s); };
        // the datum of n's unrooted parent entered from n
        auto parent_side = [&](int32_t n) -> int32_t {
            const int32_t q = parent[n];
            if (q == root) {                                     // the sibling under the root, entered from n:
                const int32_t x = q - (int32_t)n_taxa;           // equal to the sibling's post-order datum
                return ch[2 * x] == n ? ch[2 * x + 1] : ch[2 * x];
            }
            return up_ref(q, side[n]);
        };
        int32_t *he = F->h_edges.data() + (size_t)t * F->n_edges * 2;
        {
            const int32_t x = root - (int32_t)n_taxa;
            he[0] = std::min(ch[2 * x], ch[2 * x + 1]); he[1] = std::max(ch[2 * x], ch[2 * x + 1]);
        }
        uint32_t ne = 1;
        for (const int32_t n : order) {
            if ((uint32_t)n < n_taxa || n == root) continue;
            const int32_t x = n - (int32_t)n_taxa;
            const int32_t c1 = std::min(ch[2 * x], ch[2 * x + 1]), c2 = std::max(ch[2 * x], ch[2 * x + 1]);
            const int32_t ps = parent_side(n);
            const size_t d = (size_t)depth[n];
            if (by_depth.size() <= d) by_depth.resize(d + 1);
            // rotations (i, j, k) of [c1, c2, p]: entered from c1 = DO(c2-side, p-side); from c2 = DO(p-side, c1-side)
            by_depth[d].push_back({(int32_t)t, (int32_t)(ni + 2 * (uint32_t)x + 0), c2, ps});
            by_depth[d].push_back({(int32_t)t, (int32_t)(ni + 2 * (uint32_t)x + 1), ps, c1});
            // the two edges below n: DO("n entered from the child", the child's post-order datum)
            // (with kept edge data the result slot of edge e is 3 * n_internal + e)
            const int32_t eslot = F->keep_edges ? (int32_t)(3 * ni) : 0;
            edges.push_back({(int32_t)t, eslot + (int32_t)ne, up_ref(n, 0), c1}); he[2 * ne] = n; he[2 * ne + 1] = c1; ++ne;
            edges.push_back({(int32_t)t, eslot + (int32_t)ne, up_ref(n, 1), c2}); he[2 * ne] = n; he[2 * ne + 1] = c2; ++ne;
        }
        if (ne != F->n_edges) {
            ctx->err = "re-rooting: internal error enumerating the unrooted edges";
            return PCGDO_ERR_ARG;
        }
    }
    const size_t n_post = (size_t)F->n_trees * ni;
    std::vector<LevelItem> extra;
    F->up_levels.clear();
    for (auto &lv : by_depth) {
        if (lv.empty()) continue;
        F->up_levels.push_back({n_post + extra.size(), lv.size()});
        extra.insert(extra.end(), lv.begin(), lv.end());
    }
    F->edge_items = {n_post + extra.size(), edges.size()};
    extra.insert(extra.end(), edges.begin(), edges.end());
    const uint32_t stride = 3 * ni + (F->keep_edges ? F->n_edges : 0u);
    if (F->keep_edges) {
        // trial jobs: (edge datum, trial leaf) for every tree and edge; edge 0's datum is the post-order root
        F->trial_items = {n_post + extra.size(), (size_t)F->n_trees * F->n_edges};
        for (uint32_t t = 0; t < F->n_trees; ++t)
            for (uint32_t e = 0; e < F->n_edges; ++e)
                extra.push_back({(int32_t)t, (int32_t)e,
                                 e == 0 ? (int32_t)(n_taxa + (uint32_t)F->h_roots[t]) : (int32_t)(n_taxa + 3 * ni + e),
                                 (int32_t)(n_taxa + stride)});
    }
    if (n_post + extra.size() > F->items_cap) {
        ctx->err = "re-rooting: internal error sizing the job lists";
        return PCGDO_ERR_ARG;
    }
    CK(cudaMemcpy((LevelItem *)F->items.p + n_post, extra.data(), extra.size() * sizeof(LevelItem), cudaMemcpyHostToDevice));
    return PCGDO_OK;
}

// Grows the node storage to 3 result slots per internal node (+ one per edge with keep_edge_data) and builds the
// job lists.  Call once, before or after pcgdo_forest_set_leaves.
extern "C" int pcgdo_forest_enable_rerooting(pcgdo_ctx *ctx, pcgdo_forest *F, int keep_edge_data)
{
    if (!ctx || !F) return PCGDO_ERR_ARG;
    if (F->reroot) return (keep_edge_data && !F->keep_edges) ? PCGDO_ERR_ARG : PCGDO_OK;
    if (F->n_taxa < 3) {
        ctx->err = "re-rooting needs at least 3 taxa";
        return PCGDO_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    F->keep_edges = keep_edge_data != 0;
    const uint32_t ni = F->n_internal;
    int rc;
    // item array: post-order items + 2 (ni - 1) directed data + (n_edges - 1) edge jobs (+ n_edges trial jobs) per tree
    const size_t n_post = (size_t)F->n_trees * ni;
    const size_t cap = n_post + (size_t)F->n_trees * (2 * (size_t)(ni - 1) + (F->n_edges - 1) + (F->keep_edges ? F->n_edges : 0));
    {
        DevBuf items;
        if ((rc = pcgdo_ensure(ctx, items, cap * sizeof(LevelItem)))) return rc;
        CK(cudaMemcpy(items.p, F->items.p, n_post * sizeof(LevelItem), cudaMemcpyDeviceToDevice));
        CK(cudaFree(F->items.p));
        F->items = items;
        F->items_cap = cap;
    }
    const uint32_t stride = 3 * ni + (F->keep_edges ? F->n_edges : 0u);
    // 3 result slots per internal node (+ one per edge, + the trial leaf); the leaf region (front of the store) is preserved
    const size_t n_slots = (size_t)F->n_trees * F->n_chars * stride;
    DevBuf store;
    F->trial_off = F->leaf_bytes + n_slots * F->cap;
    if ((rc = pcgdo_ensure(ctx, store, (F->trial_off + (size_t)F->n_chars * F->cap + 64) * F->elem_bytes))) return rc;
    if ((rc = pcgdo_ensure(ctx, F->trial_len, (size_t)F->n_chars * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, F->trial_delta, (size_t)F->n_trees * F->n_edges * 8))) return rc;
    CK(cudaMemcpy(store.p, F->store.p, F->leaf_bytes * F->elem_bytes, cudaMemcpyDeviceToDevice));
    CK(cudaFree(F->store.p));
    F->store = store;
    if ((rc = pcgdo_ensure(ctx, F->node_len, n_slots * 4)) || (rc = pcgdo_ensure(ctx, F->node_local, n_slots * 4)) ||
        (rc = pcgdo_ensure(ctx, F->node_total, n_slots * 8)))
        return rc;
    const size_t n_cols = (size_t)F->n_trees * F->n_chars;
    if ((rc = pcgdo_ensure(ctx, F->edge_total, n_cols * F->n_edges * 8)) ||
        (rc = pcgdo_ensure(ctx, F->edge_local, n_cols * F->n_edges * 4)) || (rc = pcgdo_ensure(ctx, F->char_min, n_cols * 8)))
        return rc;
    F->stride = stride;
    F->reroot = true;
    return forest_build_reroot(ctx, F);
}

// Post-order, directed data (pre-order by depth), edge costs; per-tree cost = sum over characters of the minimum
// over the tree's unrooted edges (every dynamic character picks its own best root edge).
extern "C" int pcgdo_forest_score_rerooted(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, pcgdo_forest *F, void *cuda_stream)
{
    if (!ctx || !F || (!tcm && F->elem_bytes == 1)) return PCGDO_ERR_ARG;
    if (!F->reroot) {
        ctx->err = "call pcgdo_forest_enable_rerooting first";
        return PCGDO_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const ForestDev f = dev_view(F);
    ctx->last_launches = 0;
    CK(cudaEventRecord(ctx->ev0, st));
    bool first = true;
    LinearParams p{};
    int rc;
    for (const auto &lv : F->levels)
        if ((rc = forest_run_items(ctx, tcm, F, st, lv.first, lv.second, 0, first, p))) return rc;
    for (const auto &lv : F->up_levels)
        if ((rc = forest_run_items(ctx, tcm, F, st, lv.first, lv.second, 0, first, p))) return rc;
    if ((rc = forest_run_items(ctx, tcm, F, st, F->edge_items.first, F->edge_items.second, F->keep_edges ? 2 : 1, first, p)))
        return rc;
    const uint64_t n_cols = (uint64_t)F->n_trees * F->n_chars;
    forest_edge_min_kernel<<<(uint32_t)((n_cols + 127) / 128), 128, 0, st>>>(
        f, (const int32_t *)F->roots.p, F->n_trees, F->n_edges, (int64_t *)F->edge_total.p, (int32_t *)F->edge_local.p,
        (int64_t *)F->char_min.p);
    CK(cudaGetLastError());
    forest_reduce_min_kernel<<<(F->n_trees + 127) / 128, 128, 0, st>>>(F->n_chars, (const int64_t *)F->char_min.p,
                                                                      F->n_trees, (int64_t *)F->tree_cost.p);
    CK(cudaGetLastError());
    ctx->last_launches += 2;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    return forest_check(ctx, F, p, st);
}

// Wagner-build edge trials (app/pcg/PCG/Command/Build/Evaluate.hs:494-527, iterativeBuild.tryEdge): attaching a new leaf
// to edge e costs  cost(DO(edge datum, leaf)) - cost(edge datum) - cost(leaf) = the LOCAL cost of that alignment; the
// edge datum is the median of rooting the tree on e, kept by the re-rooting pass (keep_edge_data).  One cost-only
// batch over all trees x edges x characters.  leaf strings: HOST, character c = elems[off[c] .. off[c]+len[c]).
// delta_out: HOST int64[n_trees * n_edges], summed over this forest's characters.
extern "C" int pcgdo_forest_try_leaf(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, pcgdo_forest *F, const uint8_t *elems,
                                     const uint64_t *off, const uint32_t *len, int64_t *delta_out, void *cuda_stream)
{
    // (elements of `elems` are F's element width: bytes, or 32-bit words for a forest made by pcgdo_forest_create32)
    if (!ctx || !F || (!tcm && F->elem_bytes == 1) || !elems || !off || !len || !delta_out) return PCGDO_ERR_ARG;
    if (!F->reroot || !F->keep_edges) {
        ctx->err = "edge trials need pcgdo_forest_enable_rerooting(..., keep_edge_data = 1) and a re-rooted sweep";
        return PCGDO_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const size_t eb = F->elem_bytes;
    std::vector<uint8_t> stage((size_t)F->n_chars * F->cap * eb, 0);
    for (uint32_t c = 0; c < F->n_chars; ++c) {
        if (len[c] == PCGDO_MISSING) continue;
        if (len[c] > F->cap) {
            ctx->err = "forest: the trial leaf is longer than the node capacity";
            return PCGDO_ERR_ARG;
        }
        std::copy(elems + off[c] * eb, elems + (off[c] + len[c]) * eb, stage.begin() + (size_t)c * F->cap * eb);
    }
    CK(cudaMemcpy((uint8_t *)F->store.p + F->trial_off * eb, stage.data(), stage.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpyAsync(F->trial_len.p, len, (size_t)F->n_chars * 4, cudaMemcpyHostToDevice, st));
    ctx->last_launches = 0;
    CK(cudaEventRecord(ctx->ev0, st));
    bool first = true;
    LinearParams p{};
    int rc;
    if ((rc = forest_run_items(ctx, tcm, F, st, F->trial_items.first, F->trial_items.second, 3, first, p))) return rc;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    CK(cudaMemcpyAsync(delta_out, F->trial_delta.p, (size_t)F->n_trees * F->n_edges * 8, cudaMemcpyDeviceToHost, st));
    return forest_check(ctx, F, p, st);
}

// For parity tests / reporting: the 2*n_taxa-3 unrooted edges of one tree ((u, v) node ids; edge 0 = the root edge)
// and, for one character, the total and local cost of rooting on each of them.  Any output may be NULL.
extern "C" int pcgdo_forest_get_edges(pcgdo_ctx *ctx, pcgdo_forest *F, uint32_t tree, uint32_t c, int32_t *edge_uv,
                                      int64_t *edge_total, int32_t *edge_local)
{
    if (!ctx || !F || !F->reroot || tree >= F->n_trees || c >= F->n_chars) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (edge_uv) std::copy(F->h_edges.begin() + (size_t)tree * F->n_edges * 2,
                           F->h_edges.begin() + (size_t)(tree + 1) * F->n_edges * 2, edge_uv);
    const size_t base = ((size_t)tree * F->n_chars + c) * F->n_edges;
    if (edge_total) CK(cudaMemcpy(edge_total, (int64_t *)F->edge_total.p + base, (size_t)F->n_edges * 8, cudaMemcpyDeviceToHost));
    if (edge_local) CK(cudaMemcpy(edge_local, (int32_t *)F->edge_local.p + base, (size_t)F->n_edges * 4, cudaMemcpyDeviceToHost));
    return PCGDO_OK;
}
extern "C" uint32_t pcgdo_forest_edges(const pcgdo_forest *F) { return F ? F->n_edges : 0; }

extern "C" void *pcgdo_forest_tree_cost_ptr(pcgdo_forest *F) { return F ? F->tree_cost.p : nullptr; }

// Device-to-device copy of the int64[n_trees] cost vector (e.g. into a tensor that is then all-reduced over NCCL).
extern "C" int pcgdo_forest_copy_tree_costs(pcgdo_ctx *ctx, pcgdo_forest *F, void *dst_device, void *cuda_stream)
{
    if (!ctx || !F || !dst_device) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CK(cudaMemcpyAsync(dst_device, F->tree_cost.p, (size_t)F->n_trees * 8, cudaMemcpyDeviceToDevice, st));
    return PCGDO_OK;
}

extern "C" int pcgdo_forest_tree_costs(pcgdo_ctx *ctx, pcgdo_forest *F, int64_t *host_out)
{
    if (!ctx || !F || !host_out) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(host_out, F->tree_cost.p, (size_t)F->n_trees * 8, cudaMemcpyDeviceToHost));
    return PCGDO_OK;
}

// For parity tests / reporting: one node of one (tree, character) after a sweep.
// node is an internal node id (n_taxa + x).  median_out needs node_cap bytes.
extern "C" int pcgdo_forest_get_node(pcgdo_ctx *ctx, pcgdo_forest *F, uint32_t tree, uint32_t c, uint32_t node,
                                     uint8_t *median_out, uint32_t *len_out, int32_t *local_cost, int64_t *total_cost)
{
    if (!ctx || !F || tree >= F->n_trees || c >= F->n_chars || node < F->n_taxa || node >= 2 * F->n_taxa - 1)
        return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint32_t x = node - F->n_taxa;
    const size_t me = ((size_t)tree * F->n_chars + c) * F->stride + x;
    uint32_t len = 0;
    CK(cudaMemcpy(&len, (uint32_t *)F->node_len.p + me, 4, cudaMemcpyDeviceToHost));
    if (len_out) *len_out = len;
    if (local_cost) CK(cudaMemcpy(local_cost, (int32_t *)F->node_local.p + me, 4, cudaMemcpyDeviceToHost));
    if (total_cost) CK(cudaMemcpy(total_cost, (int64_t *)F->node_total.p + me, 8, cudaMemcpyDeviceToHost));
    if (median_out && len && len != PCGDO_MISSING)      // (len elements of F's element width)
        CK(cudaMemcpy(median_out, (uint8_t *)F->store.p + (F->leaf_bytes + me * F->cap) * F->elem_bytes,
                      (size_t)len * F->elem_bytes, cudaMemcpyDeviceToHost));
    return PCGDO_OK;
}

extern "C" uint32_t pcgdo_forest_levels(const pcgdo_forest *F) { return F ? (uint32_t)F->levels.size() : 0; }
extern "C" uint64_t pcgdo_forest_alignments(const pcgdo_forest *F)
{
    return F ? (uint64_t)F->n_trees * F->n_chars * F->n_internal : 0;
}
