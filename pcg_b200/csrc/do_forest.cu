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
#include "do_ctx.h"

#include <algorithm>
#include <cstdio>

using namespace pcgdo;

struct LevelItem { int32_t tree, x, left, right; };   // x = internal index (node id = n_taxa + x)

struct pcgdo_forest {
    uint32_t n_trees = 0, n_taxa = 0, n_chars = 0, cap = 0, n_internal = 0;
    size_t leaf_bytes = 0;
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
    size_t max_level_items = 0;
    std::vector<int32_t> h_roots;
};

namespace {

struct ForestDev {
    uint32_t n_taxa, n_chars, n_internal, cap;
    uint64_t leaf_bytes;
    const uint32_t *leaf_len;
    uint32_t *node_len;
    int32_t *node_local;
    int64_t *node_total;
};

__device__ __forceinline__ uint64_t slot_of(const ForestDev &f, uint32_t tree, uint32_t c, int32_t node)
{
    if ((uint32_t)node < f.n_taxa) return ((uint64_t)node * f.n_chars + c) * f.cap;
    return f.leaf_bytes + (((uint64_t)tree * f.n_chars + c) * f.n_internal + (uint32_t)(node - f.n_taxa)) * f.cap;
}
__device__ __forceinline__ uint64_t nidx(const ForestDev &f, uint32_t tree, uint32_t c, uint32_t x)
{
    return ((uint64_t)tree * f.n_chars + c) * f.n_internal + x;
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
    uoff[k] = slot_of(f, it.tree, c, (int32_t)(f.n_taxa + it.x));
    len1[k] = (uint32_t)it.left < f.n_taxa ? f.leaf_len[(uint64_t)it.left * f.n_chars + c]
                                           : f.node_len[nidx(f, it.tree, c, it.left - f.n_taxa)];
    len2[k] = (uint32_t)it.right < f.n_taxa ? f.leaf_len[(uint64_t)it.right * f.n_chars + c]
                                            : f.node_len[nidx(f, it.tree, c, it.right - f.n_taxa)];
}

__global__ void forest_post_kernel(ForestDev f, const LevelItem *items, uint32_t n_items, const int32_t *cost,
                                   const uint32_t *ulen)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)n_items * f.n_chars) return;
    const LevelItem it = items[k / f.n_chars];
    const uint32_t c = (uint32_t)(k % f.n_chars);
    const uint64_t me = nidx(f, it.tree, c, it.x);
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

ForestDev dev_view(const pcgdo_forest *F)
{
    ForestDev f;
    f.n_taxa = F->n_taxa; f.n_chars = F->n_chars; f.n_internal = F->n_internal; f.cap = F->cap;
    f.leaf_bytes = F->leaf_bytes;
    f.leaf_len = (const uint32_t *)F->leaf_len.p;
    f.node_len = (uint32_t *)F->node_len.p;
    f.node_local = (int32_t *)F->node_local.p;
    f.node_total = (int64_t *)F->node_total.p;
    return f;
}

}  // namespace

extern "C" int pcgdo_forest_create(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars,
                                   uint32_t node_cap, const int32_t *children, pcgdo_forest **out)
{
    if (!ctx || !children || !out || n_trees == 0 || n_taxa < 2 || n_chars == 0 || node_cap < 4) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    pcgdo_forest *F = new pcgdo_forest();
    F->n_trees = n_trees; F->n_taxa = n_taxa; F->n_chars = n_chars;
    F->cap = (node_cap + 3u) & ~3u;
    F->n_internal = n_taxa - 1;
    const uint32_t ni = F->n_internal, n_nodes = 2 * n_taxa - 1;
    // heights per tree -> items grouped by height
    std::vector<std::vector<LevelItem>> by_h;
    F->h_roots.resize(n_trees);
    std::vector<int> height(n_nodes), is_child(n_nodes);
    for (uint32_t t = 0; t < n_trees; ++t) {
        const int32_t *ch = children + (size_t)t * ni * 2;
        std::fill(height.begin(), height.end(), -1);
        std::fill(is_child.begin(), is_child.end(), 0);
        for (uint32_t v = 0; v < n_taxa; ++v) height[v] = 0;
        for (uint32_t x = 0; x < ni; ++x)
            for (int s = 0; s < 2; ++s) {
                const int32_t c = ch[2 * x + s];
                if (c < 0 || (uint32_t)c >= n_nodes || is_child[c]) {
                    ctx->err = "forest: children[] is not a binary tree over the taxa";
                    delete F;
                    return PCGDO_ERR_ARG;
                }
                is_child[c] = 1;
            }
        // iterate to a fixed point (children may be listed in any order)
        bool progress = true;
        uint32_t done = 0;
        while (progress && done < ni) {
            progress = false;
            for (uint32_t x = 0; x < ni; ++x) {
                if (height[n_taxa + x] >= 0) continue;
                const int hl = height[ch[2 * x]], hr = height[ch[2 * x + 1]];
                if (hl < 0 || hr < 0) continue;
                height[n_taxa + x] = 1 + std::max(hl, hr);
                progress = true;
                ++done;
            }
        }
        int root = -1;
        for (uint32_t x = 0; x < ni; ++x)
            if (!is_child[n_taxa + x]) root = root < 0 ? (int)x : -2;
        if (done < ni || root < 0) {
            ctx->err = "forest: children[] has a cycle or more than one root";
            delete F;
            return PCGDO_ERR_ARG;
        }
        F->h_roots[t] = root;
        for (uint32_t x = 0; x < ni; ++x) {
            const size_t h = (size_t)height[n_taxa + x];
            if (by_h.size() <= h) by_h.resize(h + 1);
            by_h[h].push_back({(int32_t)t, (int32_t)x, ch[2 * x], ch[2 * x + 1]});
        }
    }
    std::vector<LevelItem> all;
    for (size_t h = 1; h < by_h.size(); ++h) {
        if (by_h[h].empty()) continue;
        F->levels.push_back({all.size(), by_h[h].size()});
        F->max_level_items = std::max(F->max_level_items, by_h[h].size());
        all.insert(all.end(), by_h[h].begin(), by_h[h].end());
    }
    F->leaf_bytes = (size_t)n_taxa * n_chars * F->cap;
    const size_t n_slots = (size_t)n_trees * n_chars * ni;
    int rc;
    auto fail = [&](int r) { delete F; return r; };
    if ((rc = pcgdo_ensure(ctx, F->store, F->leaf_bytes + n_slots * F->cap + 64))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->leaf_len, (size_t)n_taxa * n_chars * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->node_len, n_slots * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->node_local, n_slots * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->node_total, n_slots * 8))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->items, all.size() * sizeof(LevelItem)))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->roots, (size_t)n_trees * 4))) return fail(rc);
    if ((rc = pcgdo_ensure(ctx, F->tree_cost, (size_t)n_trees * 8))) return fail(rc);
    const size_t mb = F->max_level_items * n_chars;
    if ((rc = pcgdo_ensure(ctx, F->boff1, mb * 8)) || (rc = pcgdo_ensure(ctx, F->boff2, mb * 8)) ||
        (rc = pcgdo_ensure(ctx, F->buoff, mb * 8)) || (rc = pcgdo_ensure(ctx, F->blen1, mb * 4)) ||
        (rc = pcgdo_ensure(ctx, F->blen2, mb * 4)) || (rc = pcgdo_ensure(ctx, F->bcost, mb * 4)) ||
        (rc = pcgdo_ensure(ctx, F->bulen, mb * 4)))
        return fail(rc);
    CK(cudaMemcpy(F->items.p, all.data(), all.size() * sizeof(LevelItem), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(F->roots.p, F->h_roots.data(), (size_t)n_trees * 4, cudaMemcpyHostToDevice));
    *out = F;
    return PCGDO_OK;
}

extern "C" void pcgdo_forest_destroy(pcgdo_ctx *ctx, pcgdo_forest *F)
{
    if (!F) return;
    if (ctx) cudaSetDevice(ctx->device);
    DevBuf *bufs[] = {&F->store, &F->leaf_len, &F->node_len, &F->node_local, &F->node_total, &F->items, &F->roots,
                      &F->tree_cost, &F->boff1, &F->boff2, &F->buoff, &F->blen1, &F->blen2, &F->bcost, &F->bulen};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    delete F;
}

// Leaf strings (host): character c of taxon v = elems[off[v*n_chars+c] .. +len[...]) — one byte per element,
// already free of gap elements (leaves are ungapped sequences).
extern "C" int pcgdo_forest_set_leaves(pcgdo_ctx *ctx, pcgdo_forest *F, const uint8_t *elems, const uint64_t *off,
                                       const uint32_t *len)
{
    if (!ctx || !F || !elems || !off || !len) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)F->n_taxa * F->n_chars;
    std::vector<uint8_t> stage(n * F->cap, 0);
    for (size_t i = 0; i < n; ++i) {
        if (len[i] > F->cap) {
            ctx->err = "forest: a leaf string is longer than the node capacity";
            return PCGDO_ERR_ARG;
        }
        std::copy(elems + off[i], elems + off[i] + len[i], stage.begin() + i * F->cap);
    }
    CK(cudaMemcpy(F->store.p, stage.data(), stage.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(F->leaf_len.p, len, n * 4, cudaMemcpyHostToDevice));
    return PCGDO_OK;
}

// One post-order sweep over all trees and this forest's characters.  tree_cost_out: DEVICE pointer to
// n_trees int64 (may be NULL: read the forest's own buffer with pcgdo_forest_tree_cost_ptr).
extern "C" int pcgdo_forest_score(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, pcgdo_forest *F, void *cuda_stream)
{
    if (!ctx || !tcm || !F) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const ForestDev f = dev_view(F);
    ctx->last_launches = 0;
    CK(cudaEventRecord(ctx->ev0, st));
    bool first = true;
    LinearParams p{};
    for (const auto &lv : F->levels) {
        const LevelItem *items = (const LevelItem *)F->items.p + lv.first;
        const uint32_t n_items = (uint32_t)lv.second;
        const uint64_t n_pairs = (uint64_t)n_items * F->n_chars;
        const uint32_t blocks = (uint32_t)((n_pairs + 255) / 256);
        forest_prep_kernel<<<blocks, 256, 0, st>>>(f, items, n_items, (uint64_t *)F->boff1.p, (uint64_t *)F->boff2.p,
                                                   (uint64_t *)F->buoff.p, (uint32_t *)F->blen1.p,
                                                   (uint32_t *)F->blen2.p);
        CK(cudaGetLastError());
        BatchDev bd{};
        bd.n_pairs = (uint32_t)n_pairs;
        bd.elems = (const uint8_t *)F->store.p;
        bd.off1 = (const uint64_t *)F->boff1.p; bd.off2 = (const uint64_t *)F->boff2.p;
        bd.len1 = (const uint32_t *)F->blen1.p; bd.len2 = (const uint32_t *)F->blen2.p;
        bd.cost = (int32_t *)F->bcost.p;
        bd.out_ungapped = (uint8_t *)F->store.p;
        bd.out_uoff = (const uint64_t *)F->buoff.p;
        bd.out_ulen = (uint32_t *)F->bulen.p;
        bd.ucap = F->cap;
        int rc = pcgdo_enqueue_linear(ctx, tcm, bd, st, first, &p, nullptr);
        if (rc) return rc;
        first = false;
        forest_post_kernel<<<blocks, 256, 0, st>>>(f, items, n_items, (const int32_t *)F->bcost.p,
                                                   (const uint32_t *)F->bulen.p);
        CK(cudaGetLastError());
        ctx->last_launches += 2;
    }
    forest_reduce_kernel<<<(F->n_trees + 127) / 128, 128, 0, st>>>(f, (const int32_t *)F->roots.p, F->n_trees,
                                                                  (int64_t *)F->tree_cost.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    // one look at the failure counters for the whole sweep: pairs beyond the warp kernels' limits, wide pairs
    // with the wide kernel disabled, node strings beyond their slot
    unsigned int h[10] = {0};
    CK(cudaMemcpyAsync(h, p.counter, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
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
    const size_t me = ((size_t)tree * F->n_chars + c) * F->n_internal + x;
    uint32_t len = 0;
    CK(cudaMemcpy(&len, (uint32_t *)F->node_len.p + me, 4, cudaMemcpyDeviceToHost));
    if (len_out) *len_out = len;
    if (local_cost) CK(cudaMemcpy(local_cost, (int32_t *)F->node_local.p + me, 4, cudaMemcpyDeviceToHost));
    if (total_cost) CK(cudaMemcpy(total_cost, (int64_t *)F->node_total.p + me, 8, cudaMemcpyDeviceToHost));
    if (median_out && len)
        CK(cudaMemcpy(median_out, (uint8_t *)F->store.p + F->leaf_bytes + me * F->cap, len, cudaMemcpyDeviceToHost));
    return PCGDO_OK;
}

extern "C" uint32_t pcgdo_forest_levels(const pcgdo_forest *F) { return F ? (uint32_t)F->levels.size() : 0; }
extern "C" uint64_t pcgdo_forest_alignments(const pcgdo_forest *F)
{
    return F ? (uint64_t)F->n_trees * F->n_chars * F->n_internal : 0;
}
