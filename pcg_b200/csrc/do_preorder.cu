// do_preorder.cu — pre-order pass for dynamic characters: final ("implied") alignments of every node of many rooted
// trees x characters from their post-order contexts, on the device.
//
// Reference path being replaced (one left fold over a list per tree edge and character, driven top-down):
//   lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization/Internal.hs
//     :148-177  directOptimizationPreorder, initializeRoot          (root: final alignment = its context)
//     :186-211  lexicallyDisambiguate / disambiguateElement         (single disambiguation of every element)
//     :216-233  updateFromParent                                    (Missing child inherits the parent's)
//     :236-286  deriveImpliedAlignment                              (the fold: parent's final alignment, parent's context,
//                                                                    child's context -> child's final alignment)
//   lib/core/data-structures/src/Bio/Character/Encodable/Dynamic/Element.hs:135-149  (gapElement, getContext)
//
// B200 design: the fold is two prefix sums.  Walking the parent's final alignment e_0 .. e_{n-1}:
//   * every non-Gapping e_k consumes one element of the parent's context  ->  yi(k) = #non-Gapping before k;
//   * e_k takes the child's next element x when it is an Alignment, or a Deletion (Insertion) whose parent-context
//     element ys[yi(k)] is a Deletion (Insertion) too and the child is the right (left) one  ->  xi(k) = #takes before k;
//   * out_k = e_k for a Gapping e_k, xs[xi(k)] for a take, the gap element otherwise — and the gap element for every k
//     once the child's context is used up (xi(k) >= |xs|); the reference's "Impossible happened" is xi(k) < |xs| with
//     yi(k) >= |ys|, counted in an error word.
// One warp per (tree, character, child node), 32 alignment columns per step with ballot / popc prefix counts; nodes of
// equal depth across ALL trees and characters form one launch, the tree is descended level by level with no host round
// trip; contexts and final alignments stay resident in HBM as byte planes (median, left, right[, single]).
// HBM-bound by design: per edge 3n bytes in (parent's alignment) + the two contexts + 4n bytes out.
#include <algorithm>
#include <vector>

#include "do_ctx.h"
#include "../../include/pcg_do_b200.h"

using namespace pcgdo;

struct PreItem { int32_t tree, node, parent, is_left; };

struct pcgdo_preorder {
    uint32_t n_trees = 0, n_taxa = 0, n_chars = 0, n_nodes = 0;
    std::vector<int32_t> h_roots;
    std::vector<std::pair<size_t, size_t>> levels;       // (first item, count) by depth; level 0 = the roots
    DevBuf items;
    DevBuf ctx_m, ctx_l, ctx_r, ctx_off, ctx_len;        // bytes x3, u64 / u32 [n_trees * n_chars * n_nodes]
    DevBuf ia, ia_off, ia_len, ia_missing;               // bytes, u64 / u32 [columns], u8 [columns * n_nodes]
    DevBuf err;                                          // u32 [2]: "Impossible happened" count, bad context count
    std::vector<uint64_t> h_ia_off;
    std::vector<uint32_t> h_ia_len;
    bool have_contexts = false, ran = false;
};

namespace {

constexpr uint32_t kMissingLen = PCGDO_MISSING;

struct PreDev {
    uint32_t n_chars, n_nodes;
    const uint8_t *ctx_m, *ctx_l, *ctx_r;
    const uint64_t *ctx_off;
    const uint32_t *ctx_len;
    uint8_t *ia;
    const uint64_t *ia_off;          // per column; node v, plane p at ia_off + (v * 4 + p) * round4(ia_len)
    const uint32_t *ia_len;          // per column: length of the root's context (kMissingLen: Missing)
    uint8_t *ia_missing;
    unsigned int *err;
    uint32_t gap;
    int a;
};

__device__ __forceinline__ int context_of(uint32_t l, uint32_t r) { return (l ? 2 : 0) | (r ? 1 : 0); }   // 0 G, 1 D, 2 I, 3 A

__device__ __forceinline__ uint32_t disambiguate(uint32_t m, int a, uint32_t gap)
{
    m &= (a >= 32) ? 0xffffffffu : ((1u << a) - 1u);
    if (!m) return gap;
    const int idx = min(a - 1, __ffs(m) - 1);
    return 1u << idx;
}

// level 0: the root's final alignment is its context (initializeRoot)
__global__ void preorder_root_kernel(PreDev d, const int32_t *roots, uint32_t n_trees)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_trees * d.n_chars) return;
    const uint32_t col = warp, tree = col / d.n_chars;
    const uint32_t v = (uint32_t)roots[tree];
    const uint32_t n = d.ia_len[col];
    if (n == kMissingLen) { if (lane == 0) d.ia_missing[(uint64_t)col * d.n_nodes + v] = 1; return; }
    const uint64_t co = d.ctx_off[(uint64_t)col * d.n_nodes + v];
    const uint32_t n4 = (n + 3u) & ~3u;
    uint8_t *o = d.ia + d.ia_off[col] + (uint64_t)v * 4u * n4;
    for (uint32_t k = lane; k < n; k += 32) {
        const uint32_t m = d.ctx_m[co + k];
        o[k] = (uint8_t)m;
        o[n4 + k] = d.ctx_l[co + k];
        o[2 * n4 + k] = d.ctx_r[co + k];
        o[3 * n4 + k] = (uint8_t)disambiguate(m, d.a, d.gap);
    }
    if (lane == 0) d.ia_missing[(uint64_t)col * d.n_nodes + v] = 0;
}

// one warp per (item, character): updateFromParent
__global__ void preorder_level_kernel(PreDev d, const PreItem *items, uint32_t n_items)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_items * d.n_chars) return;
    const PreItem it = items[warp / d.n_chars];
    const uint32_t c = warp % d.n_chars, col = (uint32_t)it.tree * d.n_chars + c;
    const uint64_t cn = (uint64_t)col * d.n_nodes;
    const uint32_t n = d.ia_len[col];
    if (n == kMissingLen || d.ia_missing[cn + it.parent]) {            // nothing above: Missing all the way down
        if (lane == 0) d.ia_missing[cn + it.node] = 1;
        return;
    }
    const uint32_t n4 = (n + 3u) & ~3u;
    const uint8_t *pa = d.ia + d.ia_off[col] + (uint64_t)it.parent * 4u * n4;
    uint8_t *o = d.ia + d.ia_off[col] + (uint64_t)it.node * 4u * n4;
    const uint32_t nc = d.ctx_len[cn + it.node], np = d.ctx_len[cn + it.parent];
    if (lane == 0) d.ia_missing[cn + it.node] = 0;
    if (nc == kMissingLen) {                                           // isMissing cac: the parent's, single included
        for (uint32_t k = lane; k < n4; k += 32)                       // (n4 / 4 words per plane, 4 planes)
            reinterpret_cast<uint32_t *>(o)[k] = reinterpret_cast<const uint32_t *>(pa)[k];
        return;
    }
    if (np == kMissingLen) {                                           // a non-Missing child under a Missing context
        if (lane == 0) atomicAdd(d.err + 1, 1u);
        return;
    }
    const uint64_t po = d.ctx_off[cn + it.parent], xo = d.ctx_off[cn + it.node];
    const uint8_t *pcl = d.ctx_l + po, *pcr = d.ctx_r + po;
    const uint8_t *xm = d.ctx_m + xo, *xl = d.ctx_l + xo, *xr = d.ctx_r + xo;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t carry_y = 0, carry_x = 0;
    bool bad = false;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t k = base + lane;
        const bool valid = k < n;
        uint32_t em = 0, el = 0, er = 0;
        if (valid) { em = pa[k]; el = pa[n4 + k]; er = pa[2 * n4 + k]; }
        const int cx = context_of(el, er);
        const bool non_g = valid && cx != 0;
        const unsigned by = __ballot_sync(0xffffffffu, non_g);
        const uint32_t yi = carry_y + (uint32_t)__popc(by & lt);
        int yc = 0;
        if (non_g && cx != 3 && yi < np) yc = context_of(pcl[yi], pcr[yi]);
        const bool take = non_g && (cx == 3 || (cx == 1 && yc == 1 && !it.is_left) || (cx == 2 && yc == 2 && it.is_left));
        const unsigned bx = __ballot_sync(0xffffffffu, take);
        const uint32_t xi = carry_x + (uint32_t)__popc(bx & lt);
        if (valid) {
            uint32_t om = d.gap, ol = 0, orr = 0;                      // gapElement
            if (xi < nc) {
                if (yi >= np) bad = true;                              // go (_, _, []) _ = error "Impossible happened ..."
                else if (cx == 0) { om = em; ol = el; orr = er; }
                else if (take) { om = xm[xi]; ol = xl[xi]; orr = xr[xi]; }
            }
            o[k] = (uint8_t)om;
            o[n4 + k] = (uint8_t)ol;
            o[2 * n4 + k] = (uint8_t)orr;
            o[3 * n4 + k] = (uint8_t)disambiguate(om, d.a, d.gap);
        }
        carry_y += (uint32_t)__popc(by);
        carry_x += (uint32_t)__popc(bx);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(d.err, 1u);
}

PreDev dev_view(const pcgdo_preorder *P, int a)
{
    PreDev d{};
    d.n_chars = P->n_chars; d.n_nodes = P->n_nodes;
    d.ctx_m = (const uint8_t *)P->ctx_m.p; d.ctx_l = (const uint8_t *)P->ctx_l.p; d.ctx_r = (const uint8_t *)P->ctx_r.p;
    d.ctx_off = (const uint64_t *)P->ctx_off.p; d.ctx_len = (const uint32_t *)P->ctx_len.p;
    d.ia = (uint8_t *)P->ia.p; d.ia_off = (const uint64_t *)P->ia_off.p; d.ia_len = (const uint32_t *)P->ia_len.p;
    d.ia_missing = (uint8_t *)P->ia_missing.p;
    d.err = (unsigned int *)P->err.p;
    d.a = a;
    d.gap = 1u << (a - 1);
    return d;
}

}  // namespace

// children: [n_trees][n_taxa - 1][2] (left, right) node ids — leaves 0 .. n_taxa-1, internal node x = n_taxa + x — as in
// pcgdo_forest_create.  Builds the descent order (nodes by depth) once.
extern "C" int pcgdo_preorder_create(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars,
                                     const int32_t *children, pcgdo_preorder **out)
{
    if (!ctx || !children || !out || n_trees == 0 || n_taxa < 2 || n_chars == 0) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint32_t n_int = n_taxa - 1, n_nodes = 2 * n_taxa - 1;
    auto *P = new pcgdo_preorder;
    P->n_trees = n_trees; P->n_taxa = n_taxa; P->n_chars = n_chars; P->n_nodes = n_nodes;
    P->h_roots.resize(n_trees);
    std::vector<std::vector<PreItem>> by_depth;
    std::vector<int32_t> parent(n_nodes), depth(n_nodes), order;
    for (uint32_t t = 0; t < n_trees; ++t) {
        const int32_t *ch = children + (size_t)t * n_int * 2;
        std::fill(parent.begin(), parent.end(), -1);
        for (uint32_t x = 0; x < n_int; ++x)
            for (int s = 0; s < 2; ++s) {
                const int32_t c = ch[2 * x + s];
                if (c < 0 || (uint32_t)c >= n_nodes || parent[c] != -1 || c == (int32_t)(n_taxa + x)) {
                    delete P;
                    ctx->err = "preorder: children do not describe rooted binary trees";
                    return PCGDO_ERR_ARG;
                }
                parent[c] = (int32_t)(n_taxa + x);
            }
        int32_t root = -1;
        for (uint32_t v = 0; v < n_nodes; ++v)
            if (parent[v] == -1) { if (root != -1) { root = -2; break; } root = (int32_t)v; }
        if (root < (int32_t)n_taxa) { delete P; ctx->err = "preorder: not a single rooted tree"; return PCGDO_ERR_ARG; }
        P->h_roots[t] = root;
        order.assign(1, root);
        depth[root] = 0;
        for (size_t i = 0; i < order.size(); ++i) {
            const int32_t v = order[i];
            if ((uint32_t)v < n_taxa) continue;
            for (int s = 0; s < 2; ++s) {
                const int32_t c = ch[2 * (v - (int32_t)n_taxa) + s];
                depth[c] = depth[v] + 1;
                order.push_back(c);
                if ((size_t)depth[c] > by_depth.size()) by_depth.resize(depth[c]);
                by_depth[depth[c] - 1].push_back(PreItem{(int32_t)t, c, v, s == 0 ? 1 : 0});
            }
        }
        if (order.size() != n_nodes) { delete P; ctx->err = "preorder: the tree has a cycle"; return PCGDO_ERR_ARG; }
    }
    std::vector<PreItem> flat;
    for (auto &lv : by_depth) {
        P->levels.push_back({flat.size(), lv.size()});
        flat.insert(flat.end(), lv.begin(), lv.end());
    }
    int rc;
    if ((rc = pcgdo_ensure(ctx, P->items, flat.size() * sizeof(PreItem) + 16)) ||
        (rc = pcgdo_ensure(ctx, P->err, 16)) ||
        (rc = pcgdo_ensure(ctx, P->ia_missing, (size_t)n_trees * n_chars * n_nodes))) {
        delete P;
        return rc;
    }
    CK(cudaMemcpy(P->items.p, flat.data(), flat.size() * sizeof(PreItem), cudaMemcpyHostToDevice));
    *out = P;
    return PCGDO_OK;
}

extern "C" void pcgdo_preorder_destroy(pcgdo_ctx *ctx, pcgdo_preorder *P)
{
    if (!P) return;
    if (ctx) cudaSetDevice(ctx->device);
    DevBuf *bufs[] = {&P->items, &P->ctx_m, &P->ctx_l, &P->ctx_r, &P->ctx_off, &P->ctx_len, &P->ia, &P->ia_off, &P->ia_len,
                      &P->ia_missing, &P->err};
    for (DevBuf *b : bufs) cudaFree(b->p);
    delete P;
}

// Post-order contexts of every node, HOST arrays: node v of column (tree t, character c) is the three strings
// med / left / right [off[i] .. off[i] + len[i]) with i = (t * n_chars + c) * n_nodes + v; len = PCGDO_MISSING for a
// Missing character.  A leaf's context is (s, s, s).  The final alignments' storage is sized from the roots' lengths.
extern "C" int pcgdo_preorder_set_contexts(pcgdo_ctx *ctx, pcgdo_preorder *P, const uint8_t *med, const uint8_t *left,
                                           const uint8_t *right, size_t total_bytes, const uint64_t *off, const uint32_t *len)
{
    if (!ctx || !P || !med || !left || !right || !off || !len) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t cols = (size_t)P->n_trees * P->n_chars, n = cols * P->n_nodes;
    for (size_t i = 0; i < n; ++i)
        if (len[i] != kMissingLen && off[i] + len[i] > total_bytes) {
            ctx->err = "preorder: a context lies outside the arrays";
            return PCGDO_ERR_ARG;
        }
    P->h_ia_off.assign(cols, 0);
    P->h_ia_len.assign(cols, 0);
    uint64_t at = 0;
    for (size_t col = 0; col < cols; ++col) {
        const uint32_t rl = len[col * P->n_nodes + (size_t)P->h_roots[col / P->n_chars]];
        P->h_ia_len[col] = rl;
        P->h_ia_off[col] = at;
        if (rl != kMissingLen) at += (uint64_t)P->n_nodes * 4u * ((rl + 3u) & ~3u);
    }
    int rc;
    if ((rc = pcgdo_ensure(ctx, P->ctx_m, total_bytes + 16)) || (rc = pcgdo_ensure(ctx, P->ctx_l, total_bytes + 16)) ||
        (rc = pcgdo_ensure(ctx, P->ctx_r, total_bytes + 16)) || (rc = pcgdo_ensure(ctx, P->ctx_off, n * 8)) ||
        (rc = pcgdo_ensure(ctx, P->ctx_len, n * 4)) || (rc = pcgdo_ensure(ctx, P->ia, at + 16)) ||
        (rc = pcgdo_ensure(ctx, P->ia_off, cols * 8)) || (rc = pcgdo_ensure(ctx, P->ia_len, cols * 4)))
        return rc;
    CK(cudaMemcpy(P->ctx_m.p, med, total_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(P->ctx_l.p, left, total_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(P->ctx_r.p, right, total_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(P->ctx_off.p, off, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(P->ctx_len.p, len, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(P->ia_off.p, P->h_ia_off.data(), cols * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(P->ia_len.p, P->h_ia_len.data(), cols * 4, cudaMemcpyHostToDevice));
    P->have_contexts = true;
    P->ran = false;
    return PCGDO_OK;
}

// The pre-order sweep over all trees and characters: one launch per depth.  alphabet = symbols including the gap (<= 8).
extern "C" int pcgdo_preorder_run(pcgdo_ctx *ctx, pcgdo_preorder *P, int alphabet, void *cuda_stream)
{
    if (!ctx || !P || alphabet < 2 || alphabet > 8) return PCGDO_ERR_ARG;
    if (!P->have_contexts) { ctx->err = "preorder: set the contexts first"; return PCGDO_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const PreDev d = dev_view(P, alphabet);
    ctx->last_launches = 0;
    CK(cudaEventRecord(ctx->ev0, st));
    CK(cudaMemsetAsync(P->err.p, 0, 16, st));
    // roots live in the items buffer's tail? no: a small upload of their own
    DevBuf roots;
    int rc;
    if ((rc = pcgdo_ensure(ctx, roots, (size_t)P->n_trees * 4))) return rc;
    CK(cudaMemcpyAsync(roots.p, P->h_roots.data(), (size_t)P->n_trees * 4, cudaMemcpyHostToDevice, st));
    const uint64_t root_warps = (uint64_t)P->n_trees * P->n_chars;
    preorder_root_kernel<<<(unsigned)((root_warps * 32 + 127) / 128), 128, 0, st>>>(d, (const int32_t *)roots.p, P->n_trees);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    for (const auto &lv : P->levels) {
        const uint64_t warps = (uint64_t)lv.second * P->n_chars;
        preorder_level_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(d, (const PreItem *)P->items.p + lv.first,
                                                                                     (uint32_t)lv.second);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    unsigned int h_err[2] = {0, 0};
    CK(cudaMemcpyAsync(h_err, P->err.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(roots.p);
    if (h_err[0] || h_err[1]) {
        char msg[200];
        snprintf(msg, sizeof msg, "preorder: %u edge(s) reached the reference's \"Impossible happened in "
                                  "'deriveImpliedAlignment'\", %u non-Missing node(s) under a Missing context", h_err[0], h_err[1]);
        ctx->err = msg;
        return PCGDO_ERR_OVERFLOW;
    }
    P->ran = true;
    return PCGDO_OK;
}

// Length of the final alignments of one column (the root's context length; PCGDO_MISSING when the character is Missing).
extern "C" uint32_t pcgdo_preorder_length(const pcgdo_preorder *P, uint32_t tree, uint32_t c)
{
    if (!P || tree >= P->n_trees || c >= P->n_chars || P->h_ia_len.empty()) return 0;
    return P->h_ia_len[(size_t)tree * P->n_chars + c];
}

// Final alignment of one node after pcgdo_preorder_run: median / left / right / single disambiguation, each
// pcgdo_preorder_length bytes (any may be NULL).  *missing_out = 1 for a Missing character (nothing is written).
extern "C" int pcgdo_preorder_get(pcgdo_ctx *ctx, pcgdo_preorder *P, uint32_t tree, uint32_t c, uint32_t node,
                                  uint8_t *med_out, uint8_t *left_out, uint8_t *right_out, uint8_t *single_out, int *missing_out)
{
    if (!ctx || !P || !P->ran || tree >= P->n_trees || c >= P->n_chars || node >= P->n_nodes) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t col = (size_t)tree * P->n_chars + c;
    uint8_t miss = 0;
    CK(cudaMemcpy(&miss, (uint8_t *)P->ia_missing.p + col * P->n_nodes + node, 1, cudaMemcpyDeviceToHost));
    const uint32_t n = P->h_ia_len[col];
    if (missing_out) *missing_out = (miss || n == kMissingLen) ? 1 : 0;
    if (miss || n == kMissingLen) return PCGDO_OK;
    const uint32_t n4 = (n + 3u) & ~3u;
    const uint8_t *src = (const uint8_t *)P->ia.p + P->h_ia_off[col] + (size_t)node * 4u * n4;
    uint8_t *outs[4] = {med_out, left_out, right_out, single_out};
    for (int p = 0; p < 4; ++p)
        if (outs[p]) CK(cudaMemcpy(outs[p], src + (size_t)p * n4, n, cudaMemcpyDeviceToHost));
    return PCGDO_OK;
}

// Device pointers for callers that keep working on the device: plane p (0 median, 1 left, 2 right, 3 single) of node v of
// column col starts at base + off[col] + (v * 4 + p) * round4(len[col]).
extern "C" void *pcgdo_preorder_device_ptr(pcgdo_preorder *P) { return P ? P->ia.p : nullptr; }
