// do_api.cu — C ABI of the DO backend (include/pcg_do_b200.h): context, TCM upload, batched
// launches, and the reference-compatible single-pair shims.
#include "do_ctx.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <chrono>
#include <cstdint>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

namespace pcgdo {
cudaError_t launch_linear(const LinearParams &p, bool wide, int grid, int block, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_linear_multi(const LinearParams &p, int shape, int grid, int block, size_t smem_bytes, cudaStream_t stream);
uint32_t multi_seq_cap_for(int max_len, int shape);
uint32_t seq_cap_for(int max_len, int B);
uint64_t general_words_for(uint32_t n1, uint32_t n2);
cudaError_t launch_general(const LinearParams &p, uint32_t n_list, const uint32_t *list, const uint64_t *scratch_off,
                           uint32_t *scratch, cudaStream_t stream);
}
namespace pcgdo {
struct AffParams {
    TcmDev    tcm;
    BatchDev  b;
    const uint16_t *aff_sub;
    uint32_t  gap_open;
    uint32_t *arena;
    uint32_t  arena_words;
    int       pairs_per_warp;
    unsigned int *counter;
    unsigned int *overflow_count;
    int       wave;
    unsigned int *class_count;
    unsigned int *class_take;
    uint32_t     *big_list;
    uint32_t      k_one;
};
uint32_t aff_words_for(uint32_t n1, uint32_t n2);
cudaError_t launch_affine_classify(const AffParams &p, int grid, cudaStream_t stream);
cudaError_t launch_affine(const AffParams &p, bool wide, int grid, int block, size_t smem_bytes, cudaStream_t stream);
}
#include "do_metric.cuh"
#include "do_explicit.h"
namespace pcgdo {
uint32_t metric_seq_cap(int max_len);
uint32_t metric_words_for(uint32_t n1, uint32_t n2);
cudaError_t launch_metric(const MetricParams &p, int grid, int block, size_t smem_bytes, cudaStream_t stream);
}
namespace pcgdo {
double launch_int_peak(int mode, int grid, int iters, uint32_t *scratch, cudaStream_t st);
}
using namespace pcgdo;

int pcgdo_ensure(pcgdo_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return PCGDO_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return PCGDO_OK;
}

// The experiment switches of DESIGN.md are environment variables; they are read here, once per context (and again
// when a test asks for it), so that no entry point calls getenv on its hot path.
extern "C" int pcgdo_configure_env(pcgdo_ctx *ctx)
{
    if (!ctx) return PCGDO_ERR_ARG;
    pcgdo_ctx::Knobs k;
    if (const char *e = getenv("PCGDO_HOST_CHUNK_PAIRS")) { const long v = atol(e); if (v >= 256 && v <= (1 << 20)) k.host_chunk_pairs = (size_t)v; }
    if (const char *e = getenv("PCGDO_H2D_AHEAD")) { const long v = atol(e); if (v >= 1) k.h2d_ahead = (size_t)v; }
    k.zero_copy = getenv("PCGDO_ZERO_COPY") != nullptr;
    k.trace_pipeline = getenv("PCGDO_TRACE_PIPELINE") != nullptr;
    k.no_pack16 = getenv("PCGDO_NO_PACK16") != nullptr;
    k.no_quad = getenv("PCGDO_NO_QUAD") != nullptr;
    if (const char *e = getenv("PCGDO_QUAD_FLAGS")) k.quad_flags = (unsigned)atoi(e);
    k.affine_rowwise = getenv("PCGDO_AFFINE_ROWWISE") != nullptr;
    k.metric_registers = getenv("PCGDO_METRIC_REGISTERS") != nullptr;
    k.explicit_multi_pass = getenv("PCGDO_EXPLICIT_MULTI_PASS") != nullptr;
    k.no_dict = getenv("PCGDO_NO_DICT") != nullptr;
    k.skip_ahead = getenv("PCGDO_NO_SKIP_AHEAD") ? 0 : getenv("PCGDO_SKIP_FACTOR") ? atoi(getenv("PCGDO_SKIP_FACTOR")) : 64;
    k.skip_b = getenv("PCGDO_SKIP_B") ? (unsigned)atoi(getenv("PCGDO_SKIP_B")) : 0u;
    { const char *e = getenv("PCGDO_AFFINE_VARIANT"); k.affine_variant = (e && e[0] == '1') ? 1 : 0; }
    ctx->knobs = k;
    return PCGDO_OK;
}

extern "C" int pcgdo_create(pcgdo_ctx **out, int device)
{
    if (!out) return PCGDO_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fprintf(stderr, "pcgdo_create: no usable CUDA device (%s); this library has no CPU fallback\n",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PCGDO_ERR_CUDA;
    }
    if (device < 0 || device >= n) return PCGDO_ERR_ARG;
    pcgdo_ctx *ctx = new pcgdo_ctx();
    ctx->device = device;
    auto fail = [&](const char *what, cudaError_t ee) {
        fprintf(stderr, "pcgdo_create: %s: %s\n", what, cudaGetErrorString(ee));
        delete ctx;
        return PCGDO_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
    if (prop.major < 10) {
        fprintf(stderr, "pcgdo_create: device %d is sm_%d%d; this library is built for sm_100a only\n", device,
                prop.major, prop.minor);
        delete ctx;
        return PCGDO_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return fail("event", e);
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return fail("event", e);
    pcgdo_configure_env(ctx);
    *out = ctx;
    return PCGDO_OK;
}

extern "C" void pcgdo_destroy(pcgdo_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->arena, &ctx->counters, &ctx->overflow, &ctx->big, &ctx->singles, &ctx->singles2, &ctx->phase, &ctx->gen_scratch, &ctx->gen_off, &ctx->tscratch, &ctx->d_dict, &ctx->d_elems, &ctx->d_off1, &ctx->d_off2,
                      &ctx->d_len1, &ctx->d_len2, &ctx->d_out_off, &ctx->d_og, &ctx->d_o1, &ctx->d_o2,
                      &ctx->d_olen, &ctx->d_cost, &ctx->d_cells, &ctx->d_stage, &ctx->d_actual, &ctx->d_total, &ctx->d_strag};
    if (ctx->h_total) cudaFreeHost(ctx->h_total);
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
}

// Pinned, mapped host memory for the caller's batch buffers: copies to and from it run at full PCIe rate, and the
// compact transfer writes aligned strings straight into it from the kernel (no staging).
extern "C" void *pcgdo_host_alloc(pcgdo_ctx *ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        ctx->err = std::string("pcgdo_host_alloc: ") + cudaGetErrorString(cudaGetLastError());
        return nullptr;
    }
    return p;
}

extern "C" void pcgdo_host_free(pcgdo_ctx *ctx, void *p)
{
    if (ctx) cudaSetDevice(ctx->device);
    if (p) cudaFreeHost(p);
}

extern "C" const char *pcgdo_last_error(const pcgdo_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int pcgdo_configure(pcgdo_ctx *ctx, int pairs_per_warp, int warps_per_cta, int max_len)
{
    if (!ctx) return PCGDO_ERR_ARG;
    if (pairs_per_warp) {
        if (pairs_per_warp < 1 || pairs_per_warp > 32) return PCGDO_ERR_ARG;
        ctx->pairs_per_warp = pairs_per_warp;
    }
    if (warps_per_cta) {
        if (warps_per_cta < 1 || warps_per_cta > 32) return PCGDO_ERR_ARG;
        ctx->warps_per_cta = warps_per_cta;
    }
    if (max_len) {
        if (max_len < 1) return PCGDO_ERR_ARG;
        ctx->max_len = max_len;
    }
    return PCGDO_OK;
}

// Optional finer control (not part of the reference-facing surface, used by tests / bench):
// widest diagonals-per-lane variant the shared-memory window must accommodate, arena size.
extern "C" int pcgdo_configure_ex(pcgdo_ctx *ctx, int max_b, size_t arena_words)
{
    if (!ctx) return PCGDO_ERR_ARG;
    if (max_b) {
        if (max_b != 2 && max_b != 4 && max_b != 8 && max_b != 16 && max_b != 32) return PCGDO_ERR_ARG;
        ctx->max_b = max_b;
    }
    ctx->arena_words = arena_words;
    return PCGDO_OK;
}

// --------------------------------------------------------------------------------------------- TCM

static int tcm_from_tables(pcgdo_ctx *ctx, size_t a, unsigned int gap_open, const unsigned int *cost,
                           const unsigned int *median, pcgdo_tcm **out, int affine_variant = 0)
{
    if (a < 2 || a > 8) {
        ctx->err = "alphabet size must be 2..8 for the dense-table path";
        return PCGDO_ERR_UNSUPPORTED;
    }
    CK(cudaSetDevice(ctx->device));
    const size_t dim = (size_t)1 << a;
    const unsigned int gap = 1u << (a - 1);
    auto C = [&](size_t x, size_t y) { return (long long)cost[(x << a) + y]; };
    const bool repl = a <= 5;
    const size_t rdim = repl ? 32 : dim;             // replicated layout always uses a 32 x 32 table
    std::vector<int16_t> nat(rdim * rdim, (int16_t)kBigSub);
    for (size_t l = 1; l < dim; ++l)
        for (size_t s = 1; s < dim; ++s) {
            const long long v = 4 * (C(l, s) - C(l, gap) - C(gap, s)) - 1;
            if (v < -32000 || v >= kBigSub) {
                ctx->err = "TCM costs too large for the 16-bit device table";
                return PCGDO_ERR_UNSUPPORTED;
            }
            nat[l * rdim + s] = (int16_t)v;
        }
    std::vector<int16_t> tab;
    if (repl) {
        // 32 copies, copy k entirely in shared-memory bank k: entry e of copy k at byte
        // (e >> 1) * 128 + k * 4 + (e & 1) * 2
        tab.assign(32 * 1024, 0);
        for (size_t k = 0; k < 32; ++k)
            for (size_t e = 0; e < 1024; ++e) tab[((e >> 1) * 128 + k * 4 + (e & 1) * 2) / 2] = nat[e];
    } else {
        tab = nat;
    }
    std::vector<uint8_t> med(dim * dim, 0);
    std::vector<uint32_t> cgap(dim, 0), gapc(dim, 0);
    for (size_t i = 0; i < dim * dim; ++i) med[i] = (uint8_t)median[i];
    for (size_t x = 1; x < dim; ++x) { cgap[x] = cost[(x << a) + gap]; gapc[x] = cost[((size_t)gap << a) + x]; }

    pcgdo_tcm *t = new pcgdo_tcm();
    t->a = a;
    t->gap_open = gap_open;
    auto up = [&](void **d, const void *h, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    CK(up(&t->d_sub4, tab.data(), tab.size() * 2));
    std::vector<int16_t> nat2(dim * dim, (int16_t)kBigSub);     // plain [l << a | s] for the general kernel
    for (size_t l = 1; l < dim; ++l)
        for (size_t sx = 1; sx < dim; ++sx) nat2[(l << a) + sx] = nat[l * rdim + sx];
    CK(up(&t->d_sub4_nat, nat2.data(), nat2.size() * 2));
    CK(up(&t->d_median, med.data(), med.size()));
    CK(up(&t->d_cgap, cgap.data(), cgap.size() * 4));
    CK(up(&t->d_gapc, gapc.data(), gapc.size() * 4));
    if (gap_open) {
        // substitution costs between gap-stripped elements x' = x & (gap-1) (alignCharacters.c:3636, 2770).
        // variant 0: as compiled for x86-64 — the shift count costMatrixDimension = 2^a is taken mod 32, the
        // lookup reads cost_flat[(x' << (2^a & 31)) + y']; variant 1: the patched `<< alphSize`, cost[x'][y'].
        const size_t h = dim / 2;
        std::vector<uint16_t> as(h * h, 0);
        for (size_t x = 0; x < h; ++x)
            for (size_t y = 0; y < h; ++y) {
                const size_t idx = affine_variant ? ((x << a) + y) : ((x << (dim & 31)) + y);
                const unsigned long long v = idx < dim * dim ? cost[idx] : 0ull;
                if (v > 60000ull) {
                    ctx->err = "TCM costs too large for the 16-bit affine table";
                    return PCGDO_ERR_UNSUPPORTED;
                }
                as[x * h + y] = (uint16_t)v;
            }
        CK(up(&t->d_aff_sub, as.data(), as.size() * 2));
        t->affine_variant = affine_variant;
    }
    t->dev.a = (int)a;
    t->dev.repl = repl ? 1 : 0;
    t->dev.gap = gap;
    t->dev.sub4_bytes = (uint32_t)(tab.size() * 2);
    t->dev.sub4_gapgap = (int32_t)(4 * (C(gap, gap) - C(gap, gap) - C(gap, gap)) - 1);
    {
        // packed 16-bit fill: finite cell values are re-based every 32 iterations so that their maximum returns to
        // kTop16 = 0x7000; in between a value falls at most 64 * s_minus and rises at most 64 * s_plus
        long long lo = -1, hi = -1;                       // the leading-gap entries are -1
        for (size_t l = 1; l < dim; ++l)
            for (size_t sx = 1; sx < dim; ++sx) {
                const long long v = nat[l * rdim + sx];
                lo = std::min(lo, v); hi = std::max(hi, v);
            }
        const long long s_minus = 1 - lo, s_plus = std::max(0ll, hi) + 2;
        const long long spread = 0x7000 - 64 * s_minus - 64;
        t->dev.pack_ok = (0x7000 + 64 * s_plus < 0x7800 && spread >= 8192) ? 1 : 0;
        t->dev.spread16 = t->dev.pack_ok ? (uint32_t)spread : 0u;
        if (ctx->knobs.no_pack16) t->dev.pack_ok = 0;
    }
    t->dev.sub4 = (const int16_t *)t->d_sub4;
    t->dev.sub4_nat = (const int16_t *)t->d_sub4_nat;
    t->dev.median = (const uint8_t *)t->d_median;
    t->dev.cgap = (const uint32_t *)t->d_cgap;
    t->dev.gapc = (const uint32_t *)t->d_gapc;
    *out = t;
    return PCGDO_OK;
}

extern "C" int pcgdo_tcm_create(pcgdo_ctx *ctx, const unsigned int *sigma, size_t a, unsigned int gap_open,
                                pcgdo_tcm **out)
{
    if (!ctx || !sigma || !out) return PCGDO_ERR_ARG;
    if (a < 2 || a > 8) {
        ctx->err = "alphabet size must be 2..8 for the dense-table path";
        return PCGDO_ERR_UNSUPPORTED;
    }
    const size_t dim = (size_t)1 << a;
    std::vector<unsigned int> cost(dim * dim, 0), median(dim * dim, 0);
    build_dense_tables(sigma, a, cost.data(), median.data());
    return tcm_from_tables(ctx, a, gap_open, cost.data(), median.data(), out);
}

// Affine TCM: gap_open > 0; variant 0 = the reference as compiled for x86-64, 1 = with its shift patched.
extern "C" int pcgdo_tcm_create_affine(pcgdo_ctx *ctx, const unsigned int *sigma, size_t a, unsigned int gap_open,
                                       int variant, pcgdo_tcm **out)
{
    if (!ctx || !sigma || !out || gap_open == 0) return PCGDO_ERR_ARG;
    if (a < 2 || a > 8) {
        ctx->err = "alphabet size must be 2..8 for the dense-table path";
        return PCGDO_ERR_UNSUPPORTED;
    }
    if (variant == 0 && a < 5) {
        ctx->err = "as-coded affine lookup reads out of bounds for alphabets < 5 (shift by 2^a); use variant 1";
        return PCGDO_ERR_UNSUPPORTED;
    }
    const size_t dim = (size_t)1 << a;
    std::vector<unsigned int> cost(dim * dim, 0), median(dim * dim, 0);
    build_dense_tables(sigma, a, cost.data(), median.data());
    return tcm_from_tables(ctx, a, gap_open, cost.data(), median.data(), out, variant);
}

extern "C" void pcgdo_tcm_destroy(pcgdo_ctx *ctx, pcgdo_tcm *t)
{
    if (!t) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(t->d_sub4);
    cudaFree(t->d_aff_sub);
    cudaFree(t->d_sub4_nat);
    cudaFree(t->d_median);
    cudaFree(t->d_cgap);
    cudaFree(t->d_gapc);
    delete t;
}

// --------------------------------------------------------------------------------------------- batches

static size_t smem_for(const pcgdo_tcm *tcm, int warps, uint32_t seq_cap)
{
    const size_t dim = (size_t)1 << tcm->a;
    return tcm->dev.sub4_bytes + 2 * dim * 4 + (size_t)warps * (512 + (size_t)seq_cap);
}

int pcgdo_enqueue_linear(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const BatchDev &bd, cudaStream_t st,
                         bool reset_overflow, LinearParams *out_p, bool *wide_ok_out)
{
    // shared-memory window per warp: both strings + paddings for the widest variant of each kernel
    const uint32_t cap_n = seq_cap_for(ctx->max_len, ctx->max_b < 8 ? ctx->max_b : 8);
    const uint32_t cap_w = seq_cap_for(ctx->max_len, 32);
    int warps_n = ctx->warps_per_cta, warps_w = ctx->warps_per_cta < 16 ? ctx->warps_per_cta : 16;
    while (warps_n > 1 && smem_for(tcm, warps_n, cap_n) > 227 * 1024) --warps_n;
    while (warps_w > 1 && smem_for(tcm, warps_w, cap_w) > 227 * 1024) --warps_w;
    if (smem_for(tcm, warps_n, cap_n) > 227 * 1024) {
        ctx->err = "max_len too large for the shared-memory staging window";
        return PCGDO_ERR_UNSUPPORTED;
    }
    const bool wide_ok = ctx->max_b > 8 && smem_for(tcm, warps_w, cap_w) <= 227 * 1024;
    const int grid = ctx->sm_count;
    const int warps_max = warps_n > warps_w ? warps_n : warps_w;
    size_t arena_words = ctx->arena_words;
    if (!arena_words) {
        // room for pairs_per_warp near-equal-length pairs of max_len in the narrow kernel
        arena_words = (size_t)ctx->pairs_per_warp * (((size_t)ctx->max_len + 17) / 16 * 32 * 8 + ctx->max_len / 8 + 8);
    }
    arena_words = (arena_words + 3) & ~(size_t)3;
    if (arena_words > 0xfffffff0u) arena_words = 0xfffffff0u;
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->arena, (size_t)grid * warps_max * arena_words * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->counters, 128))) return rc;
    // (growing a buffer frees the old one, which synchronises the device: callers that enqueue many batches back to
    // back — the host pipeline, the forest — size these for their largest batch before the first one)
    if ((rc = pcgdo_ensure(ctx, ctx->overflow, (size_t)bd.n_pairs * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->big, (size_t)bd.n_pairs * 4))) return rc;
    // Several pairs per warp (do_linear_multi_kernel, two shapes) first, when the TCM qualifies for the packed fill; what
    // one launch leaves over goes to the next through a list on the device, and last to the narrow kernel.
    uint32_t cap_m[4] = {0, 0, 0, 0};
    int warps_q = 0;
    if (tcm->dev.pack_ok && !ctx->knobs.no_pack16 && !ctx->knobs.no_quad && bd.n_pairs >= 4) {
        const size_t dim = (size_t)1 << tcm->a;
        const size_t fixed = tcm->dev.sub4_bytes + 2 * dim * 4;
        warps_q = ctx->warps_per_cta < 16 ? ctx->warps_per_cta : 16;
        const size_t room = (227 * 1024 - fixed) / warps_q;
        for (int shape = 0; shape < 4; ++shape) {
            const uint32_t np = shape == 0 ? 4u : (shape == 1 ? 8u : (shape == 2 ? 6u : 5u));
            const uint32_t want = (multi_seq_cap_for(ctx->max_len, shape) + 15u) & ~15u;
            if (room >= 640 + np * 256) cap_m[shape] = std::min<uint32_t>((uint32_t)((room - 640) / np) & ~15u, want);
        }
        // The 8-pairs-per-warp shape (4 lanes x 32 diagonals, bands of 65 .. 127) is built and tested but OFF unless asked
        // for (PCGDO_QUAD_FLAGS bit 5): on 256-element strings it only equals the narrow kernel's B = 4 fill at 32 warps per
        // SM (1989 vs 1999 GCUPS) and costs config 5 6 % (1468 vs 1558 tree-evals/s).
        if (!(ctx->knobs.quad_flags & 32u)) cap_m[1] = 0;
        if (ctx->knobs.quad_flags & 8u) cap_m[0] = 0;          // experiment switches: without one shape
        if (ctx->knobs.quad_flags & 16u) cap_m[2] = 0;
        if (ctx->knobs.quad_flags & 64u) cap_m[3] = 0;
    }
    const bool multi = cap_m[0] || cap_m[1] || cap_m[2] || cap_m[3];
    if (multi && ((rc = pcgdo_ensure(ctx, ctx->singles, (size_t)bd.n_pairs * 4)) ||
                  (rc = pcgdo_ensure(ctx, ctx->singles2, (size_t)bd.n_pairs * 4))))
        return rc;
    CK(cudaMemsetAsync(ctx->counters.p, 0, reset_overflow ? 64 : 16, st));

    LinearParams p{};
    p.tcm = tcm->dev;
    p.b = bd;
    p.arena = (uint32_t *)ctx->arena.p;
    p.arena_words = (uint32_t)arena_words;
    p.pairs_per_warp = ctx->pairs_per_warp;
    p.counter = (unsigned int *)ctx->counters.p;
    p.overflow_count = p.counter + 8;
    p.big_count = p.counter + 3;
    p.overflow_list = (uint32_t *)ctx->overflow.p;
    // the failure counter may accumulate over several batches (reset_overflow = false) while the list is sized for
    // one: entries beyond the list are counted, not written
    p.overflow_cap = (uint32_t)std::min<size_t>(ctx->overflow.cap / 4, 0xffffffffu);
    p.big_list = (uint32_t *)ctx->big.p;
    p.k_four = 4u;
    p.k_minus1 = 0xffffffffu;
    p.l_stride = tcm->dev.repl ? 2048u : (2u << tcm->a);
    p.k_64k = 65536u;
    p.grab = 2;

    if (multi) {
        p.quad_flags = ctx->knobs.quad_flags;
        if (ctx->knobs.quad_flags & 2u) {
            if ((rc = pcgdo_ensure(ctx, ctx->phase, 64))) return rc;
            CK(cudaMemsetAsync(ctx->phase.p, 0, 64, st));
            p.phase_cycles = (unsigned long long *)ctx->phase.p;
        }
        p.quad_arena_words = (uint32_t)std::min<size_t>((arena_words * warps_max / warps_q) & ~(size_t)3, 0xfffffff0u);
        const size_t dim = (size_t)1 << tcm->a;
        // the shape most pairs are expected to take goes first (it reads the batch itself): strings of a few hundred
        // elements have bands of 65 .. 139 diagonals, ~1 kb strings 141 .. 223.  Stage s hands pairs out through
        // counter[16 + s] and leaves what it does not take in list s % 2, counted in counter[20 + s].
        CK(cudaMemsetAsync((unsigned int *)ctx->counters.p + 16, 0, 32, st));
        const bool big_first = ctx->max_len >= 700;
        const int order[4] = {big_first ? 0 : 2, big_first ? 3 : 3, big_first ? 2 : 0, 1};
        int stage = 0;
        for (int o = 0; o < 4; ++o) {
            const int shape = order[o];
            if (!cap_m[shape]) continue;
            const uint32_t np = shape == 0 ? 4u : (shape == 1 ? 8u : (shape == 2 ? 6u : 5u));
            DevBuf &in = (stage & 1) ? ctx->singles : ctx->singles2, &outb = (stage & 1) ? ctx->singles2 : ctx->singles;
            p.multi_in_list = stage == 0 ? nullptr : (const uint32_t *)in.p;
            p.multi_in_count = stage == 0 ? nullptr : p.counter + 20 + (stage - 1);
            p.multi_counter = p.counter + 16 + stage;
            p.single_list = (uint32_t *)outb.p;
            p.single_count = p.counter + 20 + stage;
            p.seq_cap = cap_m[shape];
            CK(launch_linear_multi(p, shape, grid, warps_q * 32,
                                   tcm->dev.sub4_bytes + 2 * dim * 4 + (size_t)warps_q * (640 + np * (size_t)cap_m[shape]), st));
            ctx->last_launches += 1;
            ++stage;
        }
        // (the narrow kernel reads p.single_list / p.single_count as set by the last stage)
    }
    p.seq_cap = cap_n;
    CK(launch_linear(p, false, grid, warps_n * 32, smem_for(tcm, warps_n, cap_n), st));
    ctx->last_launches += 1;
    if (wide_ok) {
        // consumes the queue the narrow kernel built on the device; exits at once when it is empty
        p.seq_cap = cap_w;
        CK(launch_linear(p, true, grid, warps_w * 32, smem_for(tcm, warps_w, cap_w), st));
        ctx->last_launches += 1;
    }

    if (out_p) *out_p = p;
    if (wide_ok_out) *wide_ok_out = wide_ok;
    return PCGDO_OK;
}

extern "C" int pcgdo_align2d_batch_device(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b, void *cuda_stream)
{
    if (!ctx || !tcm || !b) return PCGDO_ERR_ARG;
    ctx->timed = false;
    ctx->last_launches = 0;
    if (b->n_pairs == 0) return PCGDO_OK;
    if (b->n_pairs > 0x7fffffffu || !b->elems || !b->off1 || !b->off2 || !b->len1 || !b->len2 || !b->cost) {
        ctx->err = "bad batch description";
        return PCGDO_ERR_ARG;
    }
    if ((b->out_gapped || b->out_slot1 || b->out_slot2) && !b->out_off) {
        ctx->err = "out_off required when aligned strings are requested";
        return PCGDO_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;

    BatchDev bd{};
    bd.n_pairs = (uint32_t)b->n_pairs;
    bd.elems = b->elems;
    bd.off1 = b->off1; bd.off2 = b->off2;
    bd.len1 = b->len1; bd.len2 = b->len2;
    bd.out_off = b->out_off;
    bd.out_gapped = b->out_gapped; bd.out_slot1 = b->out_slot1; bd.out_slot2 = b->out_slot2;
    bd.out_len = b->out_len;
    bd.cost = b->cost;
    bd.cells = b->cells;
    LinearParams p{};
    bool wide_ok = false;
    int rc;
    if (!ctx->hold_ev0) CK(cudaEventRecord(ctx->ev0, st));
    if ((rc = pcgdo_enqueue_linear(ctx, tcm, bd, st, true, &p, &wide_ok))) return rc;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;

    unsigned int h_cnt[9] = {0};
    CK(cudaMemcpyAsync(h_cnt, p.counter, 36, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (p.phase_cycles) {
        unsigned long long h_ph[6] = {0};
        CK(cudaMemcpy(h_ph, p.phase_cycles, 48, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[pcgdo] multi-pair kernels warp-cycles: stage %llu fill %llu trace %llu emit %llu prep %llu post %llu; singles %u\n", h_ph[0], h_ph[1],
                h_ph[2], h_ph[3], h_ph[4], h_ph[5], 0u);
    }
    // pairs the warp kernels could not hold (band wider than 1023 diagonals, strings longer than the
    // staging window, scratch larger than the arena; plus the wide queue when that kernel is disabled)
    // go through the general-geometry kernel: one CTA per pair, scratch sized from their lengths
    std::vector<uint32_t> todo;
    if (h_cnt[8]) {
        todo.resize(h_cnt[8]);
        CK(cudaMemcpyAsync(todo.data(), p.overflow_list, (size_t)h_cnt[8] * 4, cudaMemcpyDeviceToHost, st));
    }
    if (!wide_ok && h_cnt[3]) {
        const size_t o = todo.size();
        todo.resize(o + h_cnt[3]);
        CK(cudaMemcpyAsync(todo.data() + o, p.big_list, (size_t)h_cnt[3] * 4, cudaMemcpyDeviceToHost, st));
    }
    if (!todo.empty()) {
        CK(cudaStreamSynchronize(st));
        std::vector<uint32_t> l1(todo.size()), l2(todo.size());
        if (todo.size() > 256) {
            std::vector<uint32_t> a1(b->n_pairs), a2(b->n_pairs);
            CK(cudaMemcpyAsync(a1.data(), b->len1, b->n_pairs * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(a2.data(), b->len2, b->n_pairs * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (size_t i = 0; i < todo.size(); ++i) { l1[i] = a1[todo[i]]; l2[i] = a2[todo[i]]; }
        } else {
            for (size_t i = 0; i < todo.size(); ++i) {
                CK(cudaMemcpyAsync(&l1[i], b->len1 + todo[i], 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(&l2[i], b->len2 + todo[i], 4, cudaMemcpyDeviceToHost, st));
            }
            CK(cudaStreamSynchronize(st));
        }
        // bounded scratch: run the list in slices of at most ~8 GB
        size_t i0 = 0;
        while (i0 < todo.size()) {
            std::vector<uint64_t> off;
            uint64_t words = 0;
            size_t i1 = i0;
            while (i1 < todo.size()) {
                const uint64_t w = general_words_for(l1[i1], l2[i1]);
                if (i1 > i0 && (words + w) * 4 > ((uint64_t)8 << 30)) break;
                off.push_back(words);
                words += w;
                ++i1;
            }
            if ((rc = pcgdo_ensure(ctx, ctx->gen_scratch, words * 4))) return rc;
            if ((rc = pcgdo_ensure(ctx, ctx->gen_off, off.size() * 8 + (i1 - i0) * 4))) return rc;
            uint64_t *d_off = (uint64_t *)ctx->gen_off.p;
            uint32_t *d_list = (uint32_t *)(d_off + off.size());
            CK(cudaMemcpyAsync(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_list, todo.data() + i0, (i1 - i0) * 4, cudaMemcpyHostToDevice, st));
            CK(launch_general(p, (uint32_t)(i1 - i0), d_list, d_off, (uint32_t *)ctx->gen_scratch.p, st));
            ctx->last_launches += 1;
            CK(cudaStreamSynchronize(st));
            i0 = i1;
        }
        CK(cudaEventRecord(ctx->ev1, st));
    }
    return PCGDO_OK;
}

// --------------------------------------------------------------------------------------------- affine

// Enqueue the affine kernels (classify + narrow + wide) for one batch on `st`: no synchronisation.  The failure
// counter (counter[8]) is reset only when `reset_overflow`; the pipelined host path accumulates it over its chunks.
static int enqueue_affine(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const BatchDev &bd, cudaStream_t st, bool reset_overflow,
                          unsigned int **overflow_out)
{
    const size_t hdim = (size_t)1 << (tcm->a - 1);
    const int warps = ctx->warps_per_cta < 16 ? ctx->warps_per_cta : 16;
    const int wave = ctx->knobs.affine_rowwise ? 0 : 1;
    const size_t tabs = (wave ? hdim * hdim * (hdim <= 16 ? 128 : 4) : 0) + ((hdim * hdim * 2 + 15) & ~(size_t)15);
    const size_t smem = tabs + (size_t)warps * 128 * 4;
    if (smem > 227 * 1024) {
        ctx->err = "alphabet too large for the affine kernel's shared-memory table";
        return PCGDO_ERR_UNSUPPORTED;
    }
    size_t arena_words = ctx->arena_words;
    if (!arena_words) {
        const size_t per = aff_words_for((uint32_t)ctx->max_len, (uint32_t)(ctx->max_len - ctx->max_len / 20));
        arena_words = (size_t)ctx->pairs_per_warp * (per ? per : 1024);
    }
    arena_words = (arena_words + 3) & ~(size_t)3;
    if (arena_words > 0xfffffff0u) arena_words = 0xfffffff0u;
    const int grid = ctx->sm_count;
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->arena, (size_t)grid * warps * arena_words * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->counters, 128))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->big, (size_t)bd.n_pairs * 4 * 6))) return rc;
    unsigned int *counter = (unsigned int *)ctx->counters.p;
    if (reset_overflow) {
        CK(cudaMemsetAsync(counter, 0, 128, st));
    } else {                                            // everything but the failure counter at [8..15]
        CK(cudaMemsetAsync(counter, 0, 32, st));
        CK(cudaMemsetAsync(counter + 16, 0, 64, st));
    }
    AffParams p{};
    p.tcm = tcm->dev;
    p.b = bd;
    p.aff_sub = (const uint16_t *)tcm->d_aff_sub;
    p.wave = wave;
    p.gap_open = tcm->gap_open;
    p.arena = (uint32_t *)ctx->arena.p;
    p.arena_words = (uint32_t)arena_words;
    p.pairs_per_warp = ctx->pairs_per_warp;
    p.counter = counter;
    p.overflow_count = counter + 8;
    p.class_count = counter + 16;
    p.class_take = counter + 24;
    p.big_list = (uint32_t *)ctx->big.p;
    p.k_one = 1u;
    CK(launch_affine_classify(p, grid, st));
    CK(launch_affine(p, false, grid, warps * 32, smem, st));
    ctx->last_launches += 2;
    if (wave) {     // consumes the queue the first kernel built on the device; exits at once when it is empty
        const int ww = warps < 8 ? warps : 8;
        CK(launch_affine(p, true, grid, ww * 32, tabs + (size_t)ww * 128 * 4, st));
        ctx->last_launches += 1;
    }
    if (overflow_out) *overflow_out = p.overflow_count;
    return PCGDO_OK;
}

static int affine_overflow_error(pcgdo_ctx *ctx, unsigned int h_over)
{
    char msg[200];
    snprintf(msg, sizeof msg, "%u affine pair(s) exceed the kernel's limits (band wider than 1024 diagonals or "
                              "scratch larger than the arena: raise max_len / arena_words)", h_over);
    ctx->err = msg;
    return PCGDO_ERR_OVERFLOW;
}

extern "C" int pcgdo_align2d_affine_batch_device(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b,
                                                 void *cuda_stream)
{
    if (!ctx || !tcm || !b) return PCGDO_ERR_ARG;
    ctx->timed = false;
    ctx->last_launches = 0;
    if (b->n_pairs == 0) return PCGDO_OK;
    if (!tcm->d_aff_sub) {
        ctx->err = "not an affine TCM (create it with pcgdo_tcm_create_affine)";
        return PCGDO_ERR_ARG;
    }
    if (b->n_pairs > 0x7fffffffu || !b->elems || !b->off1 || !b->off2 || !b->len1 || !b->len2 || !b->cost) {
        ctx->err = "bad batch description";
        return PCGDO_ERR_ARG;
    }
    if ((b->out_gapped || b->out_slot1 || b->out_slot2) && !b->out_off) {
        ctx->err = "out_off required when aligned strings are requested";
        return PCGDO_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    BatchDev bd{};
    bd.n_pairs = (uint32_t)b->n_pairs;
    bd.elems = b->elems;
    bd.off1 = b->off1; bd.off2 = b->off2;
    bd.len1 = b->len1; bd.len2 = b->len2;
    bd.out_off = b->out_off;
    bd.out_gapped = b->out_gapped; bd.out_slot1 = b->out_slot1; bd.out_slot2 = b->out_slot2;
    bd.out_len = b->out_len;
    bd.cost = b->cost;
    bd.cells = b->cells;
    if (!ctx->hold_ev0) CK(cudaEventRecord(ctx->ev0, st));
    unsigned int *d_over = nullptr;
    int rc;
    if ((rc = enqueue_affine(ctx, tcm, bd, st, true, &d_over))) return rc;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    unsigned int h_over = 0;
    CK(cudaMemcpyAsync(&h_over, d_over, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_over) return affine_overflow_error(ctx, h_over);
    return PCGDO_OK;
}

// --------------------------------------------------------------------------------------------- large alphabets

struct pcgdo_memo;
static int table_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, int alphabet, int dialect, const pcgdo_batch32 *b,
                              void *cuda_stream);

extern "C" int pcgdo_metric_batch_device(pcgdo_ctx *ctx, int alphabet, int dialect, const pcgdo_batch32 *b,
                                         void *cuda_stream)
{
    if (!ctx || !b) return PCGDO_ERR_ARG;
    // The per-pair table kernel (do_explicit.cu) with the closed-form discrete overlap; PCGDO_METRIC_REGISTERS=1
    // selects the older kernel that evaluates the overlap per cell in registers (do_metric.cu, a <= 24).
    if (!ctx->knobs.metric_registers) return table_batch_device(ctx, nullptr, alphabet, dialect, b, cuda_stream);
    ctx->timed = false;
    ctx->last_launches = 0;
    if (b->n_pairs == 0) return PCGDO_OK;
    if (alphabet < 2 || alphabet > 24 || dialect < 0 || dialect > 2) {
        ctx->err = "metric path: alphabet must be 2..24, dialect 0..2";
        return PCGDO_ERR_UNSUPPORTED;
    }
    if (b->n_pairs > 0x7fffffffu || !b->elems || !b->off1 || !b->off2 || !b->len1 || !b->len2 || !b->cost ||
        ((b->out_med || b->out_a1 || b->out_a2) && !b->out_off)) {
        ctx->err = "bad batch description";
        return PCGDO_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const uint32_t seq_cap = metric_seq_cap(ctx->max_len);
    int warps = ctx->warps_per_cta < 16 ? ctx->warps_per_cta : 16;
    while (warps > 1 && (size_t)warps * (512 + seq_cap) > 227 * 1024) --warps;
    const size_t smem = (size_t)warps * (512 + seq_cap);
    if (smem > 227 * 1024) {
        ctx->err = "max_len too large for the metric kernel's staging window";
        return PCGDO_ERR_UNSUPPORTED;
    }
    size_t arena_words = ctx->arena_words;
    if (!arena_words) {
        const size_t per = metric_words_for((uint32_t)ctx->max_len, (uint32_t)ctx->max_len);
        arena_words = (size_t)ctx->pairs_per_warp * (per ? per : 4096);
    }
    arena_words = (arena_words + 3) & ~(size_t)3;
    if (arena_words > 0xfffffff0u) arena_words = 0xfffffff0u;
    const int grid = ctx->sm_count;
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->arena, (size_t)grid * warps * arena_words * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->counters, 64))) return rc;
    CK(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
    MetricParams p{};
    p.b.n_pairs = (uint32_t)b->n_pairs;
    p.b.elems = b->elems;
    p.b.off1 = b->off1; p.b.off2 = b->off2; p.b.len1 = b->len1; p.b.len2 = b->len2;
    p.b.out_off = b->out_off;
    p.b.out_med = b->out_med; p.b.out_a1 = b->out_a1; p.b.out_a2 = b->out_a2;
    p.b.out_len = b->out_len;
    p.b.cost = b->cost;
    p.b.cells = b->cells;
    p.b.final_offset = b->final_offset;
    p.a = alphabet;
    p.dialect = dialect;
    p.arena = (uint32_t *)ctx->arena.p;
    p.arena_words = (uint32_t)arena_words;
    p.seq_cap = seq_cap;
    p.pairs_per_warp = ctx->pairs_per_warp;
    p.counter = (unsigned int *)ctx->counters.p;
    p.overflow_count = p.counter + 8;
    p.k_four = 4u;
    p.k_minus1 = 0xffffffffu;
    CK(cudaEventRecord(ctx->ev0, st));
    CK(launch_metric(p, grid, warps * 32, smem, st));
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    ctx->last_launches = 1;
    unsigned int h_over = 0;
    CK(cudaMemcpyAsync(&h_over, p.overflow_count, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_over) {
        char msg[200];
        snprintf(msg, sizeof msg, "%u pair(s) exceed the metric kernel's limits (band wider than 1023 diagonals, strings "
                                  "longer than max_len=%d, or scratch beyond the arena)", h_over, ctx->max_len);
        ctx->err = msg;
        return PCGDO_ERR_OVERFLOW;
    }
    return PCGDO_OK;
}

namespace {
struct EventPool32 {
    std::vector<cudaEvent_t> all;
    ~EventPool32() { for (cudaEvent_t e : all) cudaEventDestroy(e); }
    cudaError_t make(cudaEvent_t *out)
    {
        const cudaError_t e = cudaEventCreateWithFlags(out, cudaEventDisableTiming);
        if (e == cudaSuccess) all.push_back(*out);
        return e;
    }
};
}  // namespace

// Host buffers in, host buffers out for the large-alphabet dialects.  Batches whose strings and outputs are laid out
// monotonically (what the packers produce) are cut into chunks: chunk c + 1 is copied in and chunk c - 1 copied out
// while chunk c is aligned (three streams; the device call of a chunk returns when its kernel is done, with the
// neighbouring copies already queued).  Everything else takes the one-shot path.
// enqueue (may be empty): the same work as device_call but without any synchronisation — failure counters accumulate
// on the device over the chunks (`first` resets them) and are looked at once, after the last chunk.  A device call that
// reads anything back per chunk would queue its few bytes behind the bulk copies of the neighbouring chunks on the same
// copy engine and hold the next launch back by milliseconds (measured: 22 ms instead of 14 per 65 536-pair chunk).
using Enqueue32 = std::function<int(const pcgdo_batch32 *, cudaStream_t, bool first, unsigned int **counters)>;
// the table kernel of one chunk, asynchronously (dialects 0..2; empty for what only the synchronous call handles)
static Enqueue32 table_enqueue32(pcgdo_ctx *ctx, const pcgdo_memo *memo, int alphabet, int dialect);

template <class DeviceCall>
static int batch32_host_impl(pcgdo_ctx *ctx, const pcgdo_batch32 *b, int words, DeviceCall device_call,
                             const Enqueue32 &enqueue = Enqueue32())
{
    if (!ctx || !b) return PCGDO_ERR_ARG;
    if (b->n_pairs == 0) return PCGDO_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t n = b->n_pairs, W = (size_t)(words > 0 ? words : 1);
    const bool want = b->out_med || b->out_a1 || b->out_a2;
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_elems, b->elems_count * 4 + 16))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_off1, n * 8)) || (rc = pcgdo_ensure(ctx, ctx->d_off2, n * 8)) ||
        (rc = pcgdo_ensure(ctx, ctx->d_len1, n * 4)) || (rc = pcgdo_ensure(ctx, ctx->d_len2, n * 4)) ||
        (rc = pcgdo_ensure(ctx, ctx->d_cost, n * 4)) || (rc = pcgdo_ensure(ctx, ctx->d_olen, n * 8)) ||
        (rc = pcgdo_ensure(ctx, ctx->d_cells, n * 8)))
        return rc;
    if (want) {
        if ((rc = pcgdo_ensure(ctx, ctx->d_out_off, n * 8)) || (rc = pcgdo_ensure(ctx, ctx->d_og, b->out_count * 4 + 16)) ||
            (rc = pcgdo_ensure(ctx, ctx->d_o1, b->out_count * 4 + 16)) || (rc = pcgdo_ensure(ctx, ctx->d_o2, b->out_count * 4 + 16)))
            return rc;
    }
    cudaStream_t st = ctx->stream;
    pcgdo_batch32 d = *b;
    d.elems = (const uint32_t *)ctx->d_elems.p;
    d.off1 = (const uint64_t *)ctx->d_off1.p; d.off2 = (const uint64_t *)ctx->d_off2.p;
    d.len1 = (const uint32_t *)ctx->d_len1.p; d.len2 = (const uint32_t *)ctx->d_len2.p;
    d.out_off = want ? (const uint64_t *)ctx->d_out_off.p : nullptr;
    d.out_med = b->out_med ? (uint32_t *)ctx->d_og.p : nullptr;
    d.out_a1 = b->out_a1 ? (uint32_t *)ctx->d_o1.p : nullptr;
    d.out_a2 = b->out_a2 ? (uint32_t *)ctx->d_o2.p : nullptr;
    d.out_len = b->out_len ? (uint32_t *)ctx->d_olen.p : nullptr;
    d.cost = (int32_t *)ctx->d_cost.p;
    d.cells = b->cells ? (uint64_t *)ctx->d_cells.p : nullptr;
    d.final_offset = b->final_offset ? (int32_t *)((uint32_t *)ctx->d_olen.p + n) : nullptr;

    // ---- pipelined path
    const size_t chunk = ctx->knobs.host_chunk_pairs ? ctx->knobs.host_chunk_pairs : 65536;
    if (n > chunk && b->out_off) {
        std::vector<size_t> cb;
        for (size_t at = 0; at < n; at += chunk) cb.push_back(at);
        cb.push_back(n);
        const size_t nch = cb.size() - 1;
        std::vector<size_t> e_lo(nch), e_hi(nch), o_lo(nch), o_hi(nch);
        bool mono = true;
        size_t prev_e = 0, prev_o = 0;
        for (size_t c = 0; c < nch && mono; ++c) {              // element / output ranges of the chunk, in ELEMENTS
            size_t lo = SIZE_MAX, hi = 0, olo = SIZE_MAX, ohi = 0;
            for (size_t k = cb[c]; k < cb[c + 1]; ++k) {
                lo = std::min(lo, (size_t)std::min(b->off1[k], b->off2[k]));
                hi = std::max(hi, (size_t)std::max(b->off1[k] + b->len1[k], b->off2[k] + b->len2[k]));
                olo = std::min(olo, (size_t)b->out_off[k]);
                ohi = std::max(ohi, (size_t)b->out_off[k] + b->len1[k] + b->len2[k]);
            }
            e_lo[c] = lo; e_hi[c] = hi; o_lo[c] = olo; o_hi[c] = ohi;
            if (lo < prev_e || hi * W > b->elems_count || (want && (olo < prev_o || ohi * W > b->out_count))) mono = false;
            prev_e = hi; prev_o = ohi;
        }
        if (mono) {
            if (!ctx->s_h2d) {
                CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
                CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
            }
            EventPool32 events;
            std::vector<cudaEvent_t> ev_in(nch), ev_done(nch);
            for (size_t c = 0; c < nch; ++c) { CK(events.make(&ev_in[c])); CK(events.make(&ev_done[c])); }
            auto copy_in = [&](size_t c) -> int {
                const size_t c0 = cb[c], k = cb[c + 1] - c0;
                cudaStream_t s = ctx->s_h2d;
                CK(cudaMemcpyAsync((uint64_t *)ctx->d_off1.p + c0, b->off1 + c0, k * 8, cudaMemcpyHostToDevice, s));
                CK(cudaMemcpyAsync((uint64_t *)ctx->d_off2.p + c0, b->off2 + c0, k * 8, cudaMemcpyHostToDevice, s));
                CK(cudaMemcpyAsync((uint32_t *)ctx->d_len1.p + c0, b->len1 + c0, k * 4, cudaMemcpyHostToDevice, s));
                CK(cudaMemcpyAsync((uint32_t *)ctx->d_len2.p + c0, b->len2 + c0, k * 4, cudaMemcpyHostToDevice, s));
                if (want) CK(cudaMemcpyAsync((uint64_t *)ctx->d_out_off.p + c0, b->out_off + c0, k * 8, cudaMemcpyHostToDevice, s));
                CK(cudaMemcpyAsync((uint32_t *)ctx->d_elems.p + e_lo[c] * W, b->elems + e_lo[c] * W, (e_hi[c] - e_lo[c]) * W * 4,
                                   cudaMemcpyHostToDevice, s));
                CK(cudaEventRecord(ev_in[c], s));
                return PCGDO_OK;
            };
            auto copy_out = [&](size_t c) -> int {
                const size_t c0 = cb[c], k = cb[c + 1] - c0, ob = (o_hi[c] - o_lo[c]) * W * 4, oo = o_lo[c] * W;
                cudaStream_t s = ctx->s_d2h;
                CK(cudaStreamWaitEvent(s, ev_done[c], 0));
                CK(cudaMemcpyAsync(b->cost + c0, d.cost + c0, k * 4, cudaMemcpyDeviceToHost, s));
                if (b->out_len) CK(cudaMemcpyAsync(b->out_len + c0, d.out_len + c0, k * 4, cudaMemcpyDeviceToHost, s));
                if (b->final_offset) CK(cudaMemcpyAsync(b->final_offset + c0, d.final_offset + c0, k * 4, cudaMemcpyDeviceToHost, s));
                if (b->cells) CK(cudaMemcpyAsync(b->cells + c0, d.cells + c0, k * 8, cudaMemcpyDeviceToHost, s));
                if (b->out_med) CK(cudaMemcpyAsync(b->out_med + oo, d.out_med + oo, ob, cudaMemcpyDeviceToHost, s));
                if (b->out_a1) CK(cudaMemcpyAsync(b->out_a1 + oo, d.out_a1 + oo, ob, cudaMemcpyDeviceToHost, s));
                if (b->out_a2) CK(cudaMemcpyAsync(b->out_a2 + oo, d.out_a2 + oo, ob, cudaMemcpyDeviceToHost, s));
                return PCGDO_OK;
            };
            if ((rc = copy_in(0))) return rc;
            int launches = 0;
            unsigned int *counters = nullptr;
            for (size_t c = 0; c < nch; ++c) {
                if (c + 1 < nch && (rc = copy_in(c + 1))) return rc;
                CK(cudaStreamWaitEvent(st, ev_in[c], 0));
                const size_t c0 = cb[c];
                pcgdo_batch32 dc = d;
                dc.n_pairs = cb[c + 1] - c0;
                dc.off1 = d.off1 + c0; dc.off2 = d.off2 + c0; dc.len1 = d.len1 + c0; dc.len2 = d.len2 + c0;
                dc.out_off = d.out_off ? d.out_off + c0 : nullptr;
                dc.out_len = d.out_len ? d.out_len + c0 : nullptr;
                dc.cost = d.cost + c0;
                dc.cells = d.cells ? d.cells + c0 : nullptr;
                dc.final_offset = d.final_offset ? d.final_offset + c0 : nullptr;
                if (enqueue) {
                    if (c == 0) { ctx->last_launches = 0; CK(cudaEventRecord(ctx->ev0, st)); }
                    if ((rc = enqueue(&dc, st, c == 0, &counters))) return rc;
                } else {
                    if ((rc = device_call(&dc, st))) return rc;      // returns when the chunk's kernels are done
                    launches += ctx->last_launches;
                }
                CK(cudaEventRecord(ev_done[c], st));
                if ((rc = copy_out(c))) return rc;
            }
            if (enqueue) {
                CK(cudaEventRecord(ctx->ev1, st));
                ctx->timed = true;
                unsigned int h_over[2] = {0, 0};
                CK(cudaMemcpyAsync(h_over, counters + 8, 8, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                CK(cudaStreamSynchronize(ctx->s_d2h));
                if (h_over[0]) {
                    char msg[260];
                    snprintf(msg, sizeof msg, "%u pair(s) exceed the table kernel's limits (band wider than 1023 diagonals, strings "
                                              "longer than max_len=%d, or scratch beyond the arena); %u of them because a folded "
                                              "cost does not fit 16 bits", h_over[0], ctx->max_len, h_over[1]);
                    ctx->err = msg;
                    return PCGDO_ERR_OVERFLOW;
                }
                return PCGDO_OK;
            }
            ctx->last_launches = launches;
            CK(cudaStreamSynchronize(ctx->s_d2h));
            return PCGDO_OK;
        }
    }

    CK(cudaMemcpyAsync(ctx->d_elems.p, b->elems, b->elems_count * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_off1.p, b->off1, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_off2.p, b->off2, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_len1.p, b->len1, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_len2.p, b->len2, n * 4, cudaMemcpyHostToDevice, st));
    if (want) CK(cudaMemcpyAsync(ctx->d_out_off.p, b->out_off, n * 8, cudaMemcpyHostToDevice, st));
    if ((rc = device_call(&d, st))) return rc;
    CK(cudaMemcpyAsync(b->cost, d.cost, n * 4, cudaMemcpyDeviceToHost, st));
    if (b->out_len) CK(cudaMemcpyAsync(b->out_len, d.out_len, n * 4, cudaMemcpyDeviceToHost, st));
    if (b->final_offset) CK(cudaMemcpyAsync(b->final_offset, d.final_offset, n * 4, cudaMemcpyDeviceToHost, st));
    if (b->cells) CK(cudaMemcpyAsync(b->cells, d.cells, n * 8, cudaMemcpyDeviceToHost, st));
    if (b->out_med) CK(cudaMemcpyAsync(b->out_med, d.out_med, b->out_count * 4, cudaMemcpyDeviceToHost, st));
    if (b->out_a1) CK(cudaMemcpyAsync(b->out_a1, d.out_a1, b->out_count * 4, cudaMemcpyDeviceToHost, st));
    if (b->out_a2) CK(cudaMemcpyAsync(b->out_a2, d.out_a2, b->out_count * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PCGDO_OK;
}

extern "C" int pcgdo_metric_batch_host(pcgdo_ctx *ctx, int alphabet, int dialect, const pcgdo_batch32 *b)
{
    Enqueue32 enq;
    if (ctx && !ctx->knobs.metric_registers) enq = table_enqueue32(ctx, nullptr, alphabet, dialect);
    return batch32_host_impl(ctx, b, (alphabet + 31) / 32, [&](const pcgdo_batch32 *d, cudaStream_t st) {
        return pcgdo_metric_batch_device(ctx, alphabet, dialect, d, st);
    }, enq);
}

// --------------------------------------------------------------------------------------------- explicit matrices
// (a = 2..32): the tcm-memo replacement.  sigma[s*a + j]: s = median symbol, j = input symbol
// (costMatrix_2d.cpp:252-267 reads tcm[fixedSymbolIndex * alphabetSize + ambiguitySymbolIndex]).

struct pcgdo_memo {
    size_t a = 0;
    int kind = 2;            // kernel numbering: 0 discrete closed form, 1 L1 closed form, 2 explicit (Sankoff)
    uint32_t delta = 0;
    std::vector<uint32_t> sigma;
    void *d_sigma = nullptr;
};

// diagnoseTcm as far as the DO dispatch needs it (lib/core/tcm/src/Data/TCM/Internal.hs:506-524, 573-590): sigma(i,j) =
// |i-j| is Additive (tested first), 0/1 is NonAdditive, anything else keeps its explicit layout.  The caller factors the
// matrix (factorTCM) before handing it over — PCG applies the weight outside the DO.
static int diagnose_structure(const unsigned int *sigma, size_t a)
{
    bool additive = true, discrete = true;
    for (size_t i = 0; i < a; ++i)
        for (size_t j = 0; j < a; ++j) {
            const unsigned int v = sigma[i * a + j];
            if (v != (unsigned int)(i > j ? i - j : j - i)) additive = false;
            if (v != (i == j ? 0u : 1u)) discrete = false;
        }
    return additive ? PCGDO_METRIC_L1 : discrete ? PCGDO_METRIC_DISCRETE : PCGDO_METRIC_EXPLICIT;
}

extern "C" int pcgdo_memo_create_metric(pcgdo_ctx *ctx, const unsigned int *sigma, size_t a, int structure, pcgdo_memo **out)
{
    if (!ctx || !out) return PCGDO_ERR_ARG;
    if (a < 2 || a > PCGDO_MAX_ALPHABET) {
        ctx->err = "large-alphabet path: alphabet size must be 2..512";
        return PCGDO_ERR_UNSUPPORTED;
    }
    if (structure == PCGDO_METRIC_AUTO) {
        if (!sigma) return PCGDO_ERR_ARG;
        structure = diagnose_structure(sigma, a);
    }
    if (structure != PCGDO_METRIC_EXPLICIT && structure != PCGDO_METRIC_DISCRETE && structure != PCGDO_METRIC_L1)
        return PCGDO_ERR_ARG;
    if (structure == PCGDO_METRIC_EXPLICIT) {
        if (!sigma) return PCGDO_ERR_ARG;
        for (size_t i = 0; i < a * a; ++i)
            if (sigma[i] > 2000u) {
                ctx->err = "explicit-matrix path: symbol-change costs above 2000 do not fit the 16-bit device table";
                return PCGDO_ERR_UNSUPPORTED;
            }
    }
    CK(cudaSetDevice(ctx->device));
    pcgdo_memo *m = new pcgdo_memo();
    m->a = a;
    m->kind = structure == PCGDO_METRIC_DISCRETE ? 0 : structure == PCGDO_METRIC_L1 ? 1 : 2;
    // Ukkonen coefficient: cheapest indel of a single non-gap symbol (UnboxedUkkonenSwapping.hs:436-459) —
    // setup-time table work on a*a numbers, like setUp2dCostMtx
    if (m->kind != 2) {
        m->delta = a >= 2 ? 1u : 0u;          // discrete: 1; L1: |(a-1) - s| is smallest (1) for s = a-2
        *out = m;
        return PCGDO_OK;
    }
    m->sigma.assign(sigma, sigma + a * a);
    uint32_t delta = 0xffffffffu;
    for (size_t s = 0; s + 1 < a; ++s) {      // overlap({s}, {gap}) = min over medians t of sigma(t,s) + sigma(t,gap)
        uint32_t c = 0xffffffffu;
        for (size_t t = 0; t < a; ++t) c = std::min(c, m->sigma[t * a + s] + m->sigma[t * a + a - 1]);
        delta = std::min(delta, c);
    }
    m->delta = delta;
    cudaError_t e = cudaMalloc(&m->d_sigma, a * a * 4);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_sigma, m->sigma.data(), a * a * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        ctx->err = std::string("pcgdo_memo_create: ") + cudaGetErrorString(e);
        delete m;
        return PCGDO_ERR_CUDA;
    }
    *out = m;
    return PCGDO_OK;
}

extern "C" int pcgdo_memo_create(pcgdo_ctx *ctx, const unsigned int *sigma, size_t a, pcgdo_memo **out)
{
    if (!ctx || !sigma || !out) return PCGDO_ERR_ARG;
    if (a < 2 || a > 32) {
        ctx->err = "explicit-matrix path: alphabet size must be 2..32";
        return PCGDO_ERR_UNSUPPORTED;
    }
    return pcgdo_memo_create_metric(ctx, sigma, a, PCGDO_METRIC_EXPLICIT, out);
}

extern "C" uint32_t pcgdo_memo_alphabet(const pcgdo_memo *m) { return m ? (uint32_t)m->a : 0u; }

extern "C" int pcgdo_memo_structure(const pcgdo_memo *m)
{
    return !m ? PCGDO_METRIC_DISCRETE : m->kind == 0 ? PCGDO_METRIC_DISCRETE : m->kind == 1 ? PCGDO_METRIC_L1 : PCGDO_METRIC_EXPLICIT;
}

extern "C" void pcgdo_memo_destroy(pcgdo_ctx *ctx, pcgdo_memo *m)
{
    if (!m) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(m->d_sigma);
    delete m;
}

extern "C" int pcgdo_overlap2d_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, size_t n, const uint32_t *e1,
                                            const uint32_t *e2, uint32_t *median, uint32_t *cost, void *cuda_stream)
{
    if (!ctx || !memo || (n && (!e1 || !e2))) return PCGDO_ERR_ARG;
    if (memo->kind != 2 || memo->a > 32) {
        ctx->err = "pcgdo_overlap2d_batch: needs an explicit matrix over at most 32 symbols";
        return PCGDO_ERR_UNSUPPORTED;
    }
    ctx->timed = false;
    ctx->last_launches = 0;
    if (n == 0) return PCGDO_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CK(cudaEventRecord(ctx->ev0, st));
    CK(launch_overlap2d((const uint32_t *)memo->d_sigma, (int)memo->a, n, e1, e2, median, cost, ctx->sm_count, st));
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    ctx->last_launches = 1;
    return PCGDO_OK;
}

extern "C" int pcgdo_overlap2d_batch_host(pcgdo_ctx *ctx, const pcgdo_memo *memo, size_t n, const uint32_t *e1,
                                          const uint32_t *e2, uint32_t *median, uint32_t *cost)
{
    if (!ctx || !memo || (n && (!e1 || !e2))) return PCGDO_ERR_ARG;
    if (n == 0) return PCGDO_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_elems, n * 8)) || (rc = pcgdo_ensure(ctx, ctx->d_og, n * 8))) return rc;
    cudaStream_t st = ctx->stream;
    uint32_t *d1 = (uint32_t *)ctx->d_elems.p, *d2 = d1 + n, *dm = (uint32_t *)ctx->d_og.p, *dc = dm + n;
    CK(cudaMemcpyAsync(d1, e1, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d2, e2, n * 4, cudaMemcpyHostToDevice, st));
    if ((rc = pcgdo_overlap2d_batch_device(ctx, memo, n, d1, d2, dm, dc, st))) return rc;
    if (median) CK(cudaMemcpyAsync(median, dm, n * 4, cudaMemcpyDeviceToHost, st));
    if (cost) CK(cudaMemcpyAsync(cost, dc, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PCGDO_OK;
}

static int table_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, int alphabet, int dialect, const pcgdo_batch32 *b,
                              void *cuda_stream);

extern "C" int pcgdo_explicit_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, int dialect, const pcgdo_batch32 *b,
                                           void *cuda_stream)
{
    if (!memo) return PCGDO_ERR_ARG;
    return table_batch_device(ctx, memo, (int)memo->a, dialect, b, cuda_stream);
}

// launch geometry of the per-pair table kernel for an alphabet of `a` symbols
struct TableCfg { uint32_t seq_cap, tab_entries; int warps; size_t smem, arena_words, ts_words; };
static int table_config(pcgdo_ctx *ctx, int a, TableCfg &c, bool dict = false)
{
    c.seq_cap = metric_seq_cap(ctx->max_len);
    // 2 KB of shared memory per warp for small per-pair tables; none next to the batch dictionary (if that turns out
    // unusable the per-pair tables live in the warp's global scratch)
    c.tab_entries = dict ? 0 : 1024;
    // a <= 32: sigma staged in shared memory; wider alphabets keep it in global memory and stage the gap element only;
    // batch-dictionary mode: one replicated table per CTA instead of the per-warp ones
    const size_t sg_bytes = (a > 32 ? 64 : (((size_t)a * a * 4 + 15) & ~(size_t)15)) + (dict ? kDictSmemBytes : 0);
    const size_t per_warp = 512 + (size_t)c.seq_cap + 2 * (size_t)c.tab_entries;
    c.warps = ctx->warps_per_cta < 16 ? ctx->warps_per_cta : 16;
    if (a > 32 && c.warps > 4) c.warps = 4;                // per-warp profile scratch grows with a * max_len
    while (c.warps > 1 && sg_bytes + (size_t)c.warps * per_warp > 227 * 1024) --c.warps;
    c.smem = sg_bytes + (size_t)c.warps * per_warp;
    if (c.smem > 227 * 1024) {
        ctx->err = "max_len too large for the explicit-matrix kernel's staging window";
        return PCGDO_ERR_UNSUPPORTED;
    }
    c.arena_words = ctx->arena_words;
    if (!c.arena_words) {
        const size_t per = metric_words_for((uint32_t)ctx->max_len, (uint32_t)ctx->max_len);
        c.arena_words = (size_t)ctx->pairs_per_warp * (per ? per : 4096);
    }
    c.arena_words = (c.arena_words + 3) & ~(size_t)3;
    if (c.arena_words > 0xfffffff0u) c.arena_words = 0xfffffff0u;
    c.ts_words = (explicit_scratch_words(ctx->max_len, a, (size_t)1 << 20) + 3) & ~(size_t)3;
    if (c.ts_words > 0xfffffff0u) {
        ctx->err = "max_len x alphabet too large for the table kernel's per-warp scratch (lower max_len)";
        return PCGDO_ERR_UNSUPPORTED;
    }
    return PCGDO_OK;
}

// memo == NULL: the discrete metric over `alphabet` symbols through the same per-pair table machinery
// Enqueue the per-pair table kernel for one batch described on the device (the forest scheduler's levels): no
// synchronisation, failure counters accumulate in counters[8..10] unless reset_overflow.
int pcgdo_enqueue_table(pcgdo_ctx *ctx, const pcgdo_memo *memo, int alphabet, int dialect, const MetricBatchDev &bd,
                        cudaStream_t st, bool reset_overflow, unsigned int **counters_out, bool try_dict)
{
    const int a = memo ? (int)memo->a : alphabet;
    if (a < 2 || a > PCGDO_MAX_ALPHABET || dialect < 0 || dialect > 2) {
        ctx->err = "table path: alphabet size must be 2..512, dialect 0..2";
        return PCGDO_ERR_UNSUPPORTED;
    }
    TableCfg c, cd;
    int rc;
    if ((rc = table_config(ctx, a, c))) return rc;
    // Batch dictionary (a <= 32): one pass over the batch's elements; when they are few (amino-acid strings with a handful
    // of ambiguity codes) a single table over them, replicated per shared-memory bank, serves every pair — see DictDev.
    // The KERNEL decides (DictDev.ok, read on the device): nothing comes back to the host, the call stays asynchronous.
    // Attempted when reserving the table's 65 KB costs at most a quarter of the resident warps.
    bool dict = try_dict && a <= 32 && !ctx->knobs.no_dict && bd.n_pairs >= 64;
    if (dict && (table_config(ctx, a, cd, true) != PCGDO_OK || 4 * cd.warps < 3 * c.warps)) dict = false;
    if (dict) c = cd;
    const uint32_t seq_cap = c.seq_cap, tab_entries = c.tab_entries;
    const int warps = c.warps, grid = ctx->sm_count;
    const size_t smem = c.smem, arena_words = c.arena_words, ts_words = c.ts_words;
    if ((rc = pcgdo_ensure(ctx, ctx->arena, (size_t)grid * warps * arena_words * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->tscratch, (size_t)grid * warps * ts_words * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->counters, 64))) return rc;
    CK(cudaMemsetAsync(ctx->counters.p, 0, reset_overflow ? 64 : 16, st));
    ExplicitParams p{};
    p.b = bd;
    p.a = a;
    p.dialect = dialect;
    p.kind = memo ? memo->kind : 0;
    p.pack16 = ctx->knobs.no_pack16 ? 0 : 1;
    p.skip_ahead = ctx->knobs.skip_ahead;
    p.sigma = memo ? (const uint32_t *)memo->d_sigma : nullptr;
    p.delta = memo ? memo->delta : 1u;
    p.k_64k = 65536u;
    p.arena = (uint32_t *)ctx->arena.p;
    p.arena_words = (uint32_t)arena_words;
    p.tscratch = (uint32_t *)ctx->tscratch.p;
    p.tscratch_words = (uint32_t)ts_words;
    p.seq_cap = seq_cap;
    p.smem_tab_entries = tab_entries;
    p.pairs_per_warp = ctx->pairs_per_warp;
    p.counter = (unsigned int *)ctx->counters.p;
    p.overflow_count = p.counter + 8;
    p.k_four = 4u;
    p.k_minus1 = 0xffffffffu;
    p.k_one = 1u;
    if (dict) {
        if ((rc = pcgdo_ensure(ctx, ctx->d_dict, sizeof(DictDev)))) return rc;
        p.dict = (DictDev *)ctx->d_dict.p;
        CK(launch_dict(p, (DictDev *)ctx->d_dict.p, grid, st));
        ctx->last_launches += 2;
    }
    CK(launch_explicit(p, grid, warps * 32, smem, st));
    ctx->last_launches += 1;
    if (counters_out) *counters_out = p.counter;
    return PCGDO_OK;
}

static int table_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, int alphabet, int dialect, const pcgdo_batch32 *b,
                              void *cuda_stream)
{
    if (!ctx || !b) return PCGDO_ERR_ARG;
    ctx->timed = false;
    ctx->last_launches = 0;
    if (b->n_pairs == 0) return PCGDO_OK;
    if (dialect < 0 || dialect > 3) {
        ctx->err = "explicit-matrix path: dialect must be 0..3";
        return PCGDO_ERR_UNSUPPORTED;
    }
    if (dialect == 3 && alphabet > 32) {
        ctx->err = "dialect 3 (unboxedFullMatrixDO) is built for alphabets of at most 32 symbols";
        return PCGDO_ERR_UNSUPPORTED;
    }
    if (b->n_pairs > 0x7fffffffu || !b->elems || !b->off1 || !b->off2 || !b->len1 || !b->len2 || !b->cost ||
        ((b->out_med || b->out_a1 || b->out_a2) && !b->out_off)) {
        ctx->err = "bad batch description";
        return PCGDO_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const int a = alphabet;
    if (a < 2 || a > PCGDO_MAX_ALPHABET) {
        ctx->err = "table path: alphabet size must be 2..512";
        return PCGDO_ERR_UNSUPPORTED;
    }
    TableCfg c;
    int rc;
    if ((rc = table_config(ctx, a, c))) return rc;
    const uint32_t seq_cap = c.seq_cap, tab_entries = c.tab_entries;
    const int warps = c.warps, grid = ctx->sm_count;
    const size_t smem = c.smem, arena_words = c.arena_words, ts_words = c.ts_words;
    if ((rc = pcgdo_ensure(ctx, ctx->arena, (size_t)grid * warps * arena_words * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->tscratch, (size_t)grid * (dialect == 3 && warps < 8 ? 8 : warps) * ts_words * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->counters, 64))) return rc;
    CK(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
    ExplicitParams p{};
    p.b.n_pairs = (uint32_t)b->n_pairs;
    p.b.elems = b->elems;
    p.b.off1 = b->off1; p.b.off2 = b->off2; p.b.len1 = b->len1; p.b.len2 = b->len2;
    p.b.out_off = b->out_off;
    p.b.out_med = b->out_med; p.b.out_a1 = b->out_a1; p.b.out_a2 = b->out_a2;
    p.b.out_len = b->out_len;
    p.b.cost = b->cost;
    p.b.cells = b->cells;
    p.b.final_offset = b->final_offset;
    p.a = a;
    p.dialect = dialect;
    p.kind = memo ? memo->kind : 0;
    p.pack16 = ctx->knobs.no_pack16 ? 0 : 1;
    p.skip_ahead = ctx->knobs.skip_ahead;
    p.skip_b = ctx->knobs.skip_b;
    p.sigma = memo ? (const uint32_t *)memo->d_sigma : nullptr;
    p.delta = memo ? memo->delta : 1u;
    p.k_64k = 65536u;
    p.arena = (uint32_t *)ctx->arena.p;
    p.arena_words = (uint32_t)arena_words;
    p.tscratch = (uint32_t *)ctx->tscratch.p;
    p.tscratch_words = (uint32_t)ts_words;
    p.seq_cap = seq_cap;
    p.smem_tab_entries = tab_entries;
    p.pairs_per_warp = ctx->pairs_per_warp;
    p.counter = (unsigned int *)ctx->counters.p;
    p.overflow_count = p.counter + 8;
    p.k_four = 4u;
    p.k_minus1 = 0xffffffffu;
    p.k_one = 1u;
    CK(cudaEventRecord(ctx->ev0, st));
    unsigned int h_left = 0;
    if (dialect == 3) {                 // unboxedFullMatrixDO: its own plain kernel (8 warps per CTA, scratch from tscratch)
        CK(launch_fullmatrix(p, grid, st));
        ctx->last_launches = 1;
    } else if (dialect == 0 || !ctx->knobs.explicit_multi_pass) {
        // batch dictionary: see pcgdo_enqueue_table (the kernel reads DictDev.ok itself)
        int launches = 1;
        TableCfg cd;
        if (a <= 32 && !ctx->knobs.no_dict && b->n_pairs >= 64 && table_config(ctx, a, cd, true) == PCGDO_OK &&
            4 * cd.warps >= 3 * warps && cd.warps <= warps) {        // (arena and scratch are sized for `warps`)
            if ((rc = pcgdo_ensure(ctx, ctx->d_dict, sizeof(DictDev)))) return rc;
            DictDev *dd = (DictDev *)ctx->d_dict.p;
            CK(launch_dict(p, dd, grid, st));
            launches += 2;
            p.dict = dd;
            p.smem_tab_entries = cd.tab_entries;                     // (the shared-memory layout of the dictionary configuration)
            CK(launch_explicit(p, grid, cd.warps * 32, cd.smem, st));
        }
        if (!p.dict) CK(launch_explicit(p, grid, warps * 32, smem, st));
        ctx->last_launches = launches;
    } else {
        // (experiment, off by default: it re-builds every pair's table per pass and measured slower than one launch)
        // one launch per offset of the doubling schedule; lists and counters live on the device, launches whose list
        // is empty return at once.  The offset at least doubles per pass and a band of max_len covers everything,
        // after which the thresholds grow with the offset while the cost does not: 2 * log2 passes are plenty.
        int npass = 8;
        for (int x = ctx->max_len; x > 0; x >>= 1) npass += 2;
        if ((rc = pcgdo_ensure(ctx, ctx->big, (size_t)b->n_pairs * 8))) return rc;
        if ((rc = pcgdo_ensure(ctx, ctx->overflow, (size_t)b->n_pairs * 8))) return rc;
        if ((rc = pcgdo_ensure(ctx, ctx->counters, 64 + 8 * (size_t)(npass + 2)))) return rc;
        CK(cudaMemsetAsync(ctx->counters.p, 0, 64 + 8 * (size_t)(npass + 2), st));
        p.counter = (unsigned int *)ctx->counters.p;
        p.overflow_count = p.counter + 8;
        unsigned int *take = p.counter + 16, *cnt = take + (npass + 2);
        uint32_t *lists[2] = {(uint32_t *)ctx->big.p, (uint32_t *)ctx->big.p + b->n_pairs};
        p.pair_off = (long long *)ctx->overflow.p;
        for (int pass = 0; pass < npass; ++pass) {
            p.counter = take + pass;
            p.in_list = pass ? lists[(pass - 1) & 1] : nullptr;
            p.in_count = pass ? cnt + (pass - 1) : nullptr;
            p.out_list = lists[pass & 1];
            p.out_count = cnt + pass;
            CK(launch_explicit(p, grid, warps * 32, smem, st));
        }
        ctx->last_launches = npass;
        CK(cudaMemcpyAsync(&h_left, cnt + (npass - 1), 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    unsigned int h_over[2] = {0, 0};
    CK(cudaMemcpyAsync(h_over, p.overflow_count, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_left) {
        ctx->err = "explicit-matrix path: the doubling schedule did not terminate within the launched passes";
        return PCGDO_ERR_OVERFLOW;
    }
    if (h_over[0]) {
        char msg[260];
        snprintf(msg, sizeof msg, "%u pair(s) exceed the explicit-matrix kernel's limits (band wider than 1023 diagonals, "
                                  "strings longer than max_len=%d, or scratch beyond the arena); %u of them because a "
                                  "folded cost does not fit 16 bits", h_over[0], ctx->max_len, h_over[1]);
        ctx->err = msg;
        return PCGDO_ERR_OVERFLOW;
    }
    return PCGDO_OK;
}

static Enqueue32 table_enqueue32(pcgdo_ctx *ctx, const pcgdo_memo *memo, int alphabet, int dialect)
{
    if (!ctx || dialect < 0 || dialect > 2 || (dialect != 0 && ctx->knobs.explicit_multi_pass)) return Enqueue32();
    return [=](const pcgdo_batch32 *d, cudaStream_t st, bool first, unsigned int **counters) -> int {
        MetricBatchDev bd{};
        bd.n_pairs = (uint32_t)d->n_pairs;
        bd.elems = d->elems;
        bd.off1 = d->off1; bd.off2 = d->off2; bd.len1 = d->len1; bd.len2 = d->len2;
        bd.out_off = d->out_off;
        bd.out_med = d->out_med; bd.out_a1 = d->out_a1; bd.out_a2 = d->out_a2;
        bd.out_len = d->out_len;
        bd.cost = d->cost;
        bd.cells = d->cells;
        bd.final_offset = d->final_offset;
        return pcgdo_enqueue_table(ctx, memo, alphabet, dialect, bd, st, first, counters, true);
    };
}

extern "C" int pcgdo_explicit_batch_host(pcgdo_ctx *ctx, const pcgdo_memo *memo, int dialect, const pcgdo_batch32 *b)
{
    if (!memo) return PCGDO_ERR_ARG;
    return batch32_host_impl(ctx, b, (int)((memo->a + 31) / 32), [&](const pcgdo_batch32 *d, cudaStream_t st) {
        return pcgdo_explicit_batch_device(ctx, memo, dialect, d, st);
    }, table_enqueue32(ctx, memo, (int)memo->a, dialect));
}

extern "C" int pcgdo_last_kernel_ms(pcgdo_ctx *ctx, float *ms, int *launches)
{
    if (!ctx) return PCGDO_ERR_ARG;
    if (launches) *launches = ctx->last_launches;
    if (ms) {
        *ms = 0.f;
        if (ctx->timed) {
            CK(cudaEventSynchronize(ctx->ev1));
            CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
        }
    }
    return PCGDO_OK;
}

// Measured INT32 throughput of this GPU (integer lane-operations per second), the roofline
// denominator bench.py reports the fill kernel against.  mode 0: add/min/logic only, 1: one VIADDMNMX (ALU pipe) +
// one IMAD (FMA pipe) per pair of instructions — the most any integer kernel can issue.  The launches are short
// (2 ms), so the GPU is first held busy for ~60 ms: measured cold (round 1) the SM clock had not ramped up yet and
// the figure came out 21 % low (28.9 T against the 36.5 T the same kernel reaches when run back to back).
extern "C" int pcgdo_int32_peak(pcgdo_ctx *ctx, int mode, double *ops_per_s, float *ms_out)
{
    if (!ctx || !ops_per_s) return PCGDO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->counters, 64))) return rc;
    const int grid = ctx->sm_count * 8;
    for (int rep = 0; rep < 30; ++rep) launch_int_peak(mode, grid, 2048, (uint32_t *)ctx->counters.p + 8, ctx->stream);
    CK(cudaStreamSynchronize(ctx->stream));
    double best = 0;
    float best_ms = 0;
    for (int rep = 0; rep < 20; ++rep) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        const double ops = launch_int_peak(mode, grid, 2048, (uint32_t *)ctx->counters.p + 8, ctx->stream);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (ops / (ms * 1e-3) > best) { best = ops / (ms * 1e-3); best_ms = ms; }
    }
    *ops_per_s = best;
    if (ms_out) *ms_out = best_ms;
    return PCGDO_OK;
}

static int batch_host_impl(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b, bool affine);

extern "C" int pcgdo_align2d_batch_host(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b)
{
    return batch_host_impl(ctx, tcm, b, false);
}

// align2dAffine semantics per pair (see do_affine.cu); host buffers in, host buffers out.
extern "C" int pcgdo_align2d_affine_batch_host(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b)
{
    return batch_host_impl(ctx, tcm, b, true);
}

namespace {
// events of one host call: destroyed on every exit path
struct EventPool {
    std::vector<cudaEvent_t> all;
    ~EventPool() { for (cudaEvent_t e : all) cudaEventDestroy(e); }
    cudaError_t make(cudaEvent_t *out, unsigned flags)
    {
        const cudaError_t e = cudaEventCreateWithFlags(out, flags);
        if (e == cudaSuccess) all.push_back(*out);
        return e;
    }
};
struct HoldEv0 {
    pcgdo_ctx *c;
    explicit HoldEv0(pcgdo_ctx *ctx) : c(ctx) { c->hold_ev0 = true; }
    ~HoldEv0() { c->hold_ev0 = false; }
};
constexpr size_t kMaxStragglers = 4096;
}  // namespace

// The pairs of a pipelined batch that only the general-geometry kernel can align (band beyond 1023 diagonals, strings
// beyond the staging window, scratch beyond the warp arena): aligned as one small batch into the full-size device
// arrays at their reserved offsets, then copied out pair by pair.  `ks` = batch-wide pair indices, ascending.
// Compact transfer: their strings go behind the packed bytes of their chunk (the chunk's reserved region has room for
// them: they never claimed a packed range), out_off_actual tells the caller where.
static int host_stragglers(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b, const pcgdo_batch &d,
                           const std::vector<uint32_t> &ks, bool compact, const std::vector<size_t> &cb,
                           const std::vector<size_t> &o_lo, const std::vector<size_t> &o_hi,
                           const std::vector<uint32_t> &gran, cudaStream_t st)
{
    const size_t m = ks.size();
    const bool want_str = b->out_gapped || b->out_slot1 || b->out_slot2;
    // descriptors: [off1 | off2 | out_off] u64, [len1 | len2 | cost | out_len] u32, [cells] u64
    std::vector<uint64_t> h64(3 * m);
    std::vector<uint32_t> h32(2 * m);
    for (size_t i = 0; i < m; ++i) {
        const uint32_t k = ks[i];
        h64[i] = b->off1[k]; h64[m + i] = b->off2[k]; h64[2 * m + i] = want_str ? b->out_off[k] : 0;
        h32[i] = b->len1[k]; h32[m + i] = b->len2[k];
    }
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_strag, m * (3 * 8 + 4 * 4 + 8) + 64))) return rc;
    uint64_t *d64 = (uint64_t *)ctx->d_strag.p;
    uint64_t *d_cells = d64 + 3 * m;
    uint32_t *d32 = (uint32_t *)(d_cells + m);
    CK(cudaMemcpyAsync(d64, h64.data(), 3 * m * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d32, h32.data(), 2 * m * 4, cudaMemcpyHostToDevice, st));
    pcgdo_batch sub = d;
    sub.n_pairs = m;
    sub.off1 = d64; sub.off2 = d64 + m; sub.out_off = want_str ? d64 + 2 * m : nullptr;
    sub.len1 = d32; sub.len2 = d32 + m;
    sub.cost = (int32_t *)(d32 + 2 * m);
    sub.out_len = d32 + 3 * m;
    sub.cells = b->cells ? d_cells : nullptr;
    sub.out_off_actual = nullptr;
    const int launches = ctx->last_launches;
    {
        HoldEv0 hold(ctx);
        if ((rc = pcgdo_align2d_batch_device(ctx, tcm, &sub, st))) return rc;
    }
    ctx->last_launches += launches;
    std::vector<uint32_t> r32(2 * m);
    std::vector<uint64_t> rcells(b->cells ? m : 0);
    CK(cudaMemcpyAsync(r32.data(), d32 + 2 * m, 2 * m * 4, cudaMemcpyDeviceToHost, st));
    if (b->cells) CK(cudaMemcpyAsync(rcells.data(), d_cells, m * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    size_t c = 0, tail = 0;
    bool fresh = true;
    for (size_t i = 0; i < m; ++i) {
        const uint32_t k = ks[i];
        b->cost[k] = (int32_t)r32[i];
        if (b->out_len) b->out_len[k] = r32[m + i];
        if (b->cells) b->cells[k] = rcells[i];
        if (!want_str) continue;
        const size_t from = b->out_off[k], bytes = ((size_t)r32[m + i] + 3) & ~(size_t)3;
        size_t to = from;
        if (compact) {
            while (k >= cb[c + 1]) { ++c; fresh = true; }
            if (fresh) {
                tail = o_lo[c] + (((size_t)ctx->h_total[c] + gran[c]) & ~(size_t)gran[c]);
                fresh = false;
            }
            to = tail;
            tail += (bytes + gran[c]) & ~(size_t)gran[c];
            if (tail > o_hi[c]) {
                ctx->err = "compact transfer: a chunk's reserved output region is smaller than its alignments (overlapping out_off?)";
                return PCGDO_ERR_ARG;
            }
            b->out_off_actual[k] = to;
        }
        if (b->out_gapped) CK(cudaMemcpyAsync(b->out_gapped + to, d.out_gapped + from, bytes, cudaMemcpyDeviceToHost, st));
        if (b->out_slot1) CK(cudaMemcpyAsync(b->out_slot1 + to, d.out_slot1 + from, bytes, cudaMemcpyDeviceToHost, st));
        if (b->out_slot2) CK(cudaMemcpyAsync(b->out_slot2 + to, d.out_slot2 + from, bytes, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return PCGDO_OK;
}

static int batch_host_impl(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b, bool affine)
{
    if (!ctx || !tcm || !b) return PCGDO_ERR_ARG;
    if (b->n_pairs == 0) return PCGDO_OK;
    if (affine && !tcm->d_aff_sub) {
        ctx->err = "not an affine TCM (create it with pcgdo_tcm_create_affine)";
        return PCGDO_ERR_ARG;
    }
    if (b->n_pairs > 0x7fffffffu || !b->elems || !b->off1 || !b->off2 || !b->len1 || !b->len2 || !b->cost) {
        ctx->err = "bad batch description";
        return PCGDO_ERR_ARG;
    }
    const auto t_entry = std::chrono::steady_clock::now();
    auto ms_since_entry = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(); };
    CK(cudaSetDevice(ctx->device));
    const size_t n = b->n_pairs;
    const bool want_str = b->out_gapped || b->out_slot1 || b->out_slot2;
    if (want_str && !b->out_off) {
        ctx->err = "out_off required when aligned strings are requested";
        return PCGDO_ERR_ARG;
    }
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_elems, b->elems_bytes + 16))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_off1, n * 8))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_off2, n * 8))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_len1, n * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_len2, n * 4))) return rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_cost, n * 4))) return rc;
    if (b->out_len && (rc = pcgdo_ensure(ctx, ctx->d_olen, n * 4))) return rc;
    if (b->cells && (rc = pcgdo_ensure(ctx, ctx->d_cells, n * 8))) return rc;
    if (want_str) {
        if ((rc = pcgdo_ensure(ctx, ctx->d_out_off, n * 8))) return rc;
        if (b->out_gapped && (rc = pcgdo_ensure(ctx, ctx->d_og, b->out_bytes + 16))) return rc;
        if (b->out_slot1 && (rc = pcgdo_ensure(ctx, ctx->d_o1, b->out_bytes + 16))) return rc;
        if (b->out_slot2 && (rc = pcgdo_ensure(ctx, ctx->d_o2, b->out_bytes + 16))) return rc;
    }
    cudaStream_t st = ctx->stream;
    // descriptors (40 bytes per pair) of pairs [lo, hi): all at once on the plain path, chunk by chunk ahead of each
    // chunk's strings on the pipelined one (the first launch then waits for 0.7 MB, not 40 MB)
    auto copy_desc = [&](size_t lo, size_t hi, cudaStream_t s) -> int {
        const size_t k = hi - lo;
        CK(cudaMemcpyAsync((uint64_t *)ctx->d_off1.p + lo, b->off1 + lo, k * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync((uint64_t *)ctx->d_off2.p + lo, b->off2 + lo, k * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync((uint32_t *)ctx->d_len1.p + lo, b->len1 + lo, k * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync((uint32_t *)ctx->d_len2.p + lo, b->len2 + lo, k * 4, cudaMemcpyHostToDevice, s));
        if (want_str) CK(cudaMemcpyAsync((uint64_t *)ctx->d_out_off.p + lo, b->out_off + lo, k * 8, cudaMemcpyHostToDevice, s));
        return PCGDO_OK;
    };

    pcgdo_batch d = *b;
    d.elems = (const uint8_t *)ctx->d_elems.p;
    d.off1 = (const uint64_t *)ctx->d_off1.p;
    d.off2 = (const uint64_t *)ctx->d_off2.p;
    d.len1 = (const uint32_t *)ctx->d_len1.p;
    d.len2 = (const uint32_t *)ctx->d_len2.p;
    d.out_off = want_str ? (const uint64_t *)ctx->d_out_off.p : nullptr;
    d.out_gapped = b->out_gapped ? (uint8_t *)ctx->d_og.p : nullptr;
    d.out_slot1 = b->out_slot1 ? (uint8_t *)ctx->d_o1.p : nullptr;
    d.out_slot2 = b->out_slot2 ? (uint8_t *)ctx->d_o2.p : nullptr;
    d.out_len = b->out_len ? (uint32_t *)ctx->d_olen.p : nullptr;
    d.cost = (int32_t *)ctx->d_cost.p;
    d.cells = b->cells ? (uint64_t *)ctx->d_cells.p : nullptr;
    d.out_off_actual = nullptr;
    // ---- pipelined path: the batch is cut into chunks whose strings and outputs are contiguous ranges of
    // the host buffers; chunk c+1 is copied in and chunk c-1 copied out while chunk c is aligned
    // (three streams, PCIe is full duplex).  Needs monotone layouts, which the packers produce; anything
    // else takes the plain path below.  Affine pairs are three orders of magnitude heavier per byte (a 5 kb pair is a
    // million cells and occupies a warp for milliseconds) and their kernels come in two launches per batch (narrow band
    // classes, then wide ones): every chunk pays the end of both — warps idle while the last pairs finish.  Measured on
    // config 3 (100 000 pairs, 341 ms of kernel time): 25 chunks of 4096 pairs 1275 ms, 4 chunks of 9 .. 38 k pairs
    // 484 ms, while copy-in + one launch + copy-out is ~420 ms.  So affine batches are pipelined only when they are
    // large enough for chunks of 64 pairs per resident warp.
    size_t chunk = ctx->knobs.host_chunk_pairs ? ctx->knobs.host_chunk_pairs
                   : ctx->host_chunk_pairs     ? ctx->host_chunk_pairs : 65536;
    if (affine && !ctx->knobs.host_chunk_pairs && !ctx->host_chunk_pairs) {
        const size_t warps = (size_t)ctx->sm_count * (size_t)std::min(ctx->warps_per_cta, 16);
        chunk = std::max<size_t>(64 * warps, 8192);
    }
    bool pipelined = n > chunk && (affine || ctx->max_b > 8);     // (wide kernel on: no per-chunk queue left behind)
    // COMPACT TRANSFER (b->out_off_actual).  The reference hands every alignment back in buffers of capacity
    // len1 + len2 + 2 (DO/Pairwise/FFI.hsc:474-481), of which an alignment of two similar strings fills about half;
    // copying capacity-sized slots back made the PCIe link, not the aligner, the end-to-end bound.  Here the warps of a
    // chunk claim packed output ranges from a per-chunk cursor once a pair's alignment length is known (pack_claim):
    //  * staged (default): into one of kRing device slots, 16-byte units, copied out chunk by chunk; the host learns
    //    the byte count from a pinned word once the chunk is done.  Measured on B200 / PCIe 5: 127 -> 92 ms per 2^20
    //    pairs of 1 kb (the aligner alone: 72.6 ms);
    //  * zero-copy (PCGDO_ZERO_COPY=1, pinned mapped out arrays): the emit stores go straight into the caller's arrays
    //    over PCIe, one whole 128-byte line per warp store.  No staging and no drain, but SM-issued posted writes reached
    //    only ~23 GB/s against ~52 GB/s for the copy engine and stalled the aligner (158 ms): kept as an experiment.
    const bool compact = pipelined && !affine && want_str && b->out_off_actual && b->out_len;
    uint8_t *zc_out[3] = {nullptr, nullptr, nullptr};
    bool zero_copy = compact && ctx->knobs.zero_copy;
    if (zero_copy) {
        uint8_t *host[3] = {b->out_gapped, b->out_slot1, b->out_slot2};
        for (int a = 0; a < 3 && zero_copy; ++a) {
            if (!host[a]) continue;
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, host[a]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
                cudaGetLastError();
                zero_copy = false;
            } else {
                zc_out[a] = (uint8_t *)at.devicePointer;
            }
        }
    }
    std::vector<size_t> cb;                                       // chunk c = pairs [cb[c], cb[c+1])
    if (pipelined) {
        if (zero_copy) {
            size_t sz = std::max<size_t>(chunk / 4, 256), at = 0;
            while (at < n) { cb.push_back(at); at += sz; if (sz < 4 * chunk) sz *= 2; }
        } else {
            // a short first chunk gets the aligner going early; the copy-out lags by one chunk whatever the sizes
            // (its rate is within 10 % of the aligner's), so the rest stay equal.  (Tried in round 2: chunks of twice the
            // size in the middle and a ramp down at the end, so that a warp of the multi-pair kernel comes to its trace +
            // emit phase with a full batch — 82 ms instead of 75: a 131 072-pair chunk aligned in 9 – 10 ms next to the
            // copies against 7.5 ms alone.)
            size_t at = 0;
            for (size_t sz = std::max<size_t>(chunk / 4, 256); at < n && sz < chunk; sz *= 2) { cb.push_back(at); at += sz; }
            for (; at < n; at += chunk) cb.push_back(at);
        }
        cb.push_back(n);
    }
    const size_t nch = pipelined ? cb.size() - 1 : 0;
    std::vector<size_t> e_lo(nch), e_hi(nch), o_lo(nch), o_hi(nch);
    std::vector<uint32_t> gran(nch, 3u);                          // packing granularity - 1 of each chunk
    size_t max_chunk_pairs = 0;
    if (pipelined) {
        // per-chunk byte ranges of the strings and of the reserved outputs (a pass over 40 bytes of descriptors per
        // pair: spread over a few threads, it is on the critical path before the first copy)
        auto range_of = [&](size_t c) {
            size_t lo = SIZE_MAX, hi = 0, olo = SIZE_MAX, ohi = 0, worst16 = 0;
            for (size_t k = cb[c]; k < cb[c + 1]; ++k) {
                const size_t a1 = b->off1[k], a2 = b->off2[k];
                lo = std::min(lo, std::min(a1, a2));
                hi = std::max(hi, std::max(a1 + b->len1[k], a2 + b->len2[k]));
                if (want_str) {
                    const size_t o = b->out_off[k], cap = (size_t)b->len1[k] + b->len2[k];
                    olo = std::min(olo, o);
                    ohi = std::max(ohi, o + ((cap + 3) & ~(size_t)3));
                    worst16 += (cap + 15) & ~(size_t)15;
                }
            }
            e_lo[c] = lo; e_hi[c] = hi; o_lo[c] = olo; o_hi[c] = ohi;
            // 16-byte units (vector-friendly staging) only where even the longest possible alignments fit the
            // chunk's reserved region that way; chunks of very short strings pack in 4-byte units, which always fit
            gran[c] = (want_str && worst16 <= ohi - olo) ? 15u : 3u;
        };
        const size_t nthr = std::min<size_t>(8, nch);
        std::vector<std::thread> pool;
        for (size_t t = 1; t < nthr; ++t)
            pool.emplace_back([&, t] { for (size_t c = t; c < nch; c += nthr) range_of(c); });
        for (size_t c = 0; c < nch; c += nthr) range_of(c);
        for (auto &th : pool) th.join();
        size_t prev_e = 0, prev_o = 0;
        for (size_t c = 0; c < nch; ++c) {
            if (e_lo[c] < prev_e || e_hi[c] > b->elems_bytes || (want_str && (o_lo[c] < prev_o || o_hi[c] > b->out_bytes))) pipelined = false;
            prev_e = e_hi[c]; prev_o = want_str ? o_hi[c] : 0;
            max_chunk_pairs = std::max(max_chunk_pairs, cb[c + 1] - cb[c]);
        }
    }
    if (pipelined) {
        if (!ctx->s_h2d) {
            CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        }
        constexpr size_t kRing = 3;
        size_t slot_bytes = 0;
        if (compact) {
            if (!zero_copy) {
                for (size_t c = 0; c < nch; ++c) slot_bytes = std::max(slot_bytes, o_hi[c] - o_lo[c]);
                slot_bytes = (slot_bytes + 255) & ~(size_t)255;
                if ((rc = pcgdo_ensure(ctx, ctx->d_stage, kRing * 3 * slot_bytes))) return rc;
                if (ctx->h_total_cap < nch) {
                    if (ctx->h_total) cudaFreeHost(ctx->h_total);
                    ctx->h_total = nullptr;
                    ctx->h_total_cap = 0;
                    CK(cudaHostAlloc((void **)&ctx->h_total, nch * 8, cudaHostAllocDefault));
                    ctx->h_total_cap = nch;
                }
            }
            if ((rc = pcgdo_ensure(ctx, ctx->d_actual, n * 8)) || (rc = pcgdo_ensure(ctx, ctx->d_total, nch * 8 + 8))) return rc;
            CK(cudaMemsetAsync(ctx->d_total.p, 0, nch * 8 + 8, st));
        }
        // scratch of the kernels sized once for the largest chunk: growing it mid-pipeline would free the old
        // buffer (a device-wide synchronisation); the failure list holds batch-wide pair indices of ALL chunks
        if ((rc = pcgdo_ensure(ctx, ctx->overflow, std::max(max_chunk_pairs, std::min(n, 4 * kMaxStragglers)) * 4))) return rc;
        if ((rc = pcgdo_ensure(ctx, ctx->big, max_chunk_pairs * 4 * (affine ? 6 : 1)))) return rc;
        if (!affine && ((rc = pcgdo_ensure(ctx, ctx->singles, max_chunk_pairs * 4)) || (rc = pcgdo_ensure(ctx, ctx->singles2, max_chunk_pairs * 4))))
            return rc;
        EventPool events;
        std::vector<cudaEvent_t> ev_in(nch), ev_done(nch), ev_out(nch);
        const bool trace = ctx->knobs.trace_pipeline;             // per-chunk completion times to stderr
        const unsigned evf = trace ? cudaEventDefault : cudaEventDisableTiming;
        for (size_t c = 0; c < nch; ++c) {
            CK(events.make(&ev_in[c], evf));
            CK(events.make(&ev_done[c], evf));
            CK(events.make(&ev_out[c], evf));
        }
        const size_t h2d_ahead = ctx->knobs.h2d_ahead ? ctx->knobs.h2d_ahead : nch;
        cudaEvent_t ev_desc;
        CK(events.make(&ev_desc, cudaEventDisableTiming));
        CK(cudaEventRecord(ev_desc, st));
        CK(cudaStreamWaitEvent(ctx->s_h2d, ev_desc, 0));
        ctx->last_launches = 0;
        CK(cudaEventRecord(ctx->ev0, st));
        const double t_setup = ms_since_entry();
        LinearParams p{};
        bool wide_ok = affine;
        unsigned int *d_fail = nullptr;                           // the kernels' failure counter (accumulates over the chunks)
        unsigned int *pack_overflow = compact ? (unsigned int *)((unsigned long long *)ctx->d_total.p + nch) : nullptr;
        auto copy_in = [&](size_t c) -> int {
            int r = copy_desc(cb[c], cb[c + 1], ctx->s_h2d);
            if (r) return r;
            CK(cudaMemcpyAsync((uint8_t *)ctx->d_elems.p + e_lo[c], b->elems + e_lo[c], e_hi[c] - e_lo[c],
                               cudaMemcpyHostToDevice, ctx->s_h2d));
            CK(cudaEventRecord(ev_in[c], ctx->s_h2d));
            return PCGDO_OK;
        };
        auto copy_out = [&](size_t c) -> int {
            const size_t c0 = cb[c], c1 = cb[c + 1];
            cudaStream_t so = ctx->s_d2h;
            if (compact && !zero_copy) CK(cudaEventSynchronize(ev_done[c]));           // h_total[c] is valid from here on
            else CK(cudaStreamWaitEvent(so, ev_done[c], 0));
            CK(cudaMemcpyAsync(b->cost + c0, d.cost + c0, (c1 - c0) * 4, cudaMemcpyDeviceToHost, so));
            if (b->out_len) CK(cudaMemcpyAsync(b->out_len + c0, d.out_len + c0, (c1 - c0) * 4, cudaMemcpyDeviceToHost, so));
            if (b->cells) CK(cudaMemcpyAsync(b->cells + c0, d.cells + c0, (c1 - c0) * 8, cudaMemcpyDeviceToHost, so));
            if (compact)
                CK(cudaMemcpyAsync(b->out_off_actual + c0, (uint64_t *)ctx->d_actual.p + c0, (c1 - c0) * 8, cudaMemcpyDeviceToHost, so));
            if (compact && zero_copy) {
                // the strings are already in the caller's buffers
            } else if (compact) {
                // the kernel never claims beyond the chunk's reserved region (pack_limit), so neither does this copy
                const size_t tot = std::min((size_t)ctx->h_total[c], o_hi[c] - o_lo[c]);
                uint8_t *slot = (uint8_t *)ctx->d_stage.p + (c % kRing) * 3 * slot_bytes;
                if (b->out_gapped) CK(cudaMemcpyAsync(b->out_gapped + o_lo[c], slot, tot, cudaMemcpyDeviceToHost, so));
                if (b->out_slot1) CK(cudaMemcpyAsync(b->out_slot1 + o_lo[c], slot + slot_bytes, tot, cudaMemcpyDeviceToHost, so));
                if (b->out_slot2) CK(cudaMemcpyAsync(b->out_slot2 + o_lo[c], slot + 2 * slot_bytes, tot, cudaMemcpyDeviceToHost, so));
            } else {
                const size_t ob = want_str ? o_hi[c] - o_lo[c] : 0;
                if (b->out_gapped) CK(cudaMemcpyAsync(b->out_gapped + o_lo[c], d.out_gapped + o_lo[c], ob, cudaMemcpyDeviceToHost, so));
                if (b->out_slot1) CK(cudaMemcpyAsync(b->out_slot1 + o_lo[c], d.out_slot1 + o_lo[c], ob, cudaMemcpyDeviceToHost, so));
                if (b->out_slot2) CK(cudaMemcpyAsync(b->out_slot2 + o_lo[c], d.out_slot2 + o_lo[c], ob, cudaMemcpyDeviceToHost, so));
            }
            CK(cudaEventRecord(ev_out[c], so));
            return PCGDO_OK;
        };
        // the strings of the first chunks are queued on the copy-in stream right away, the rest h2d_ahead chunks ahead
        for (size_t c = 0; c < nch && c < h2d_ahead; ++c)
            if ((rc = copy_in(c))) return rc;
        for (size_t c = 0; c < nch; ++c) {
            const size_t c0 = cb[c], c1 = cb[c + 1];
            if (c + h2d_ahead < nch && (rc = copy_in(c + h2d_ahead))) return rc;
            CK(cudaStreamWaitEvent(st, ev_in[c], 0));
            BatchDev bd{};
            bd.n_pairs = (uint32_t)(c1 - c0);
            bd.pair_base = (uint32_t)c0;
            bd.elems = d.elems;
            bd.off1 = d.off1 + c0; bd.off2 = d.off2 + c0;
            bd.len1 = d.len1 + c0; bd.len2 = d.len2 + c0;
            bd.out_off = d.out_off ? d.out_off + c0 : nullptr;
            bd.out_gapped = d.out_gapped; bd.out_slot1 = d.out_slot1; bd.out_slot2 = d.out_slot2;
            bd.out_len = d.out_len ? d.out_len + c0 : nullptr;
            bd.cost = d.cost + c0;
            bd.cells = d.cells ? d.cells + c0 : nullptr;
            if (compact) {
                bd.pack_cursor = (unsigned long long *)ctx->d_total.p + c;
                bd.out_off_actual = (uint64_t *)ctx->d_actual.p + c0;
                bd.pack_overflow = pack_overflow;
                bd.pack_base = o_lo[c];
                bd.pack_limit = o_hi[c] - o_lo[c];          // never beyond the chunk's reserved region of the caller's arrays
                bd.pack_gran = gran[c];
                if (zero_copy) {
                    // inside the chunk's reserved region of the caller's arrays; emit_pair aligns its stores to
                    // 128-byte lines by itself
                    bd.out_gapped = zc_out[0] ? zc_out[0] + o_lo[c] : nullptr;
                    bd.out_slot1 = zc_out[1] ? zc_out[1] + o_lo[c] : nullptr;
                    bd.out_slot2 = zc_out[2] ? zc_out[2] + o_lo[c] : nullptr;
                    bd.pack_gran = 3;
                } else {
                    if (c >= kRing) CK(cudaStreamWaitEvent(st, ev_out[c - kRing], 0));      // the staging slot is free again
                    uint8_t *slot = (uint8_t *)ctx->d_stage.p + (c % kRing) * 3 * slot_bytes;
                    bd.out_gapped = b->out_gapped ? slot : nullptr;
                    bd.out_slot1 = b->out_slot1 ? slot + slot_bytes : nullptr;
                    bd.out_slot2 = b->out_slot2 ? slot + 2 * slot_bytes : nullptr;
                }
            }
            if (affine) rc = enqueue_affine(ctx, tcm, bd, st, c == 0, &d_fail);
            else {
                rc = pcgdo_enqueue_linear(ctx, tcm, bd, st, c == 0, &p, &wide_ok);
                d_fail = p.overflow_count;
            }
            if (rc) return rc;
            if (compact && !zero_copy)
                CK(cudaMemcpyAsync(ctx->h_total + c, (unsigned long long *)ctx->d_total.p + c, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ev_done[c], st));
            // chunk c-1 goes out while chunk c is aligned (staged compact mode blocks the host on chunk c-1 here, with
            // chunk c already queued behind it)
            if (c >= 1 && (rc = copy_out(c - 1))) return rc;
        }
        if ((rc = copy_out(nch - 1))) return rc;
        const double t_issued = ms_since_entry();
        CK(cudaEventRecord(ctx->ev1, st));
        ctx->timed = true;
        unsigned int h_fail = 0, h_pack = 0;
        CK(cudaMemcpyAsync(&h_fail, d_fail, 4, cudaMemcpyDeviceToHost, st));
        if (compact) CK(cudaMemcpyAsync(&h_pack, pack_overflow, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaStreamSynchronize(ctx->s_d2h));
        if (trace) {
            fprintf(stderr, "pcgdo pipeline: %zu chunks (first %zu pairs), compact=%d zero_copy=%d (ms after the first launch: in / aligned / out)\n",
                    nch, cb[1] - cb[0], (int)compact, (int)zero_copy);
            fprintf(stderr, "  host, ms after entry: set up %.2f, all issued %.2f, all done %.2f\n", t_setup, t_issued, ms_since_entry());
            for (size_t c = 0; c < nch; ++c) {
                float a = 0, d2 = 0, o = 0;
                cudaEventElapsedTime(&a, ctx->ev0, ev_in[c]);
                cudaEventElapsedTime(&d2, ctx->ev0, ev_done[c]);
                cudaEventElapsedTime(&o, ctx->ev0, ev_out[c]);
                fprintf(stderr, "  chunk %3zu  %8zu pairs  %8.2f %8.2f %8.2f\n", c, cb[c + 1] - cb[c], a, d2, o);
            }
        }
        if (affine) return h_fail ? affine_overflow_error(ctx, h_fail) : PCGDO_OK;
        // with the wide kernel enabled every pair that fits a warp has been handled in its own chunk
        if (h_fail == 0 && wide_ok && h_pack == 0) return PCGDO_OK;
        // a few pairs need the general-geometry kernel: align just those (their batch-wide indices are on the failure list)
        if (wide_ok && h_pack == 0 && !zero_copy && h_fail <= kMaxStragglers && h_fail <= p.overflow_cap) {
            std::vector<uint32_t> ks(h_fail);
            CK(cudaMemcpyAsync(ks.data(), p.overflow_list, (size_t)h_fail * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            std::sort(ks.begin(), ks.end());
            rc = host_stragglers(ctx, tcm, b, d, ks, compact, cb, o_lo, o_hi, gran, st);
            if (rc == PCGDO_OK) CK(cudaEventRecord(ctx->ev1, st));
            return rc;
        }
        // otherwise (many such pairs, or a packed range that did not fit its chunk): redo the batch on the plain path
        CK(cudaMemcpyAsync(ctx->d_elems.p, b->elems, b->elems_bytes, cudaMemcpyHostToDevice, st));   // (all of it: chunks may leave gaps)
    } else {
        if ((rc = copy_desc(0, n, st))) return rc;
        CK(cudaMemcpyAsync(ctx->d_elems.p, b->elems, b->elems_bytes, cudaMemcpyHostToDevice, st));
    }

    rc = affine ? pcgdo_align2d_affine_batch_device(ctx, tcm, &d, st) : pcgdo_align2d_batch_device(ctx, tcm, &d, st);
    if (rc) return rc;

    CK(cudaMemcpyAsync(b->cost, d.cost, n * 4, cudaMemcpyDeviceToHost, st));
    if (b->out_len) CK(cudaMemcpyAsync(b->out_len, d.out_len, n * 4, cudaMemcpyDeviceToHost, st));
    if (b->cells) CK(cudaMemcpyAsync(b->cells, d.cells, n * 8, cudaMemcpyDeviceToHost, st));
    if (b->out_gapped) CK(cudaMemcpyAsync(b->out_gapped, d.out_gapped, b->out_bytes, cudaMemcpyDeviceToHost, st));
    if (b->out_slot1) CK(cudaMemcpyAsync(b->out_slot1, d.out_slot1, b->out_bytes, cudaMemcpyDeviceToHost, st));
    if (b->out_slot2) CK(cudaMemcpyAsync(b->out_slot2, d.out_slot2, b->out_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (b->out_off_actual && b->out_off) std::copy(b->out_off, b->out_off + n, b->out_off_actual);   // nothing was packed
    return PCGDO_OK;
}

// --------------------------------------------------------------------------------------------- shims
// Reference-compatible entry points: one pair / one lookup per call through the GPU.  GHC calls them concurrently
// from its capability threads (`parZipWith rpar` over `unsafePerformIO`), so every calling thread gets its
// own small context (one stream, one warp's worth of scratch) and the calls of different threads overlap on the
// device; the only shared state is the cache of device TCMs, locked for the lookup alone.

extern "C" int pcgdo_overlap_sets_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, int nsets, size_t n, const uint32_t *e1,
                                               const uint32_t *e2, const uint32_t *e3, uint32_t *median, uint32_t *cost,
                                               void *cuda_stream)
{
    if (!ctx || !memo || (nsets != 2 && nsets != 3) || (n && (!e1 || !e2 || (nsets == 3 && !e3)))) return PCGDO_ERR_ARG;
    if (memo->kind != 2) {
        ctx->err = "pcgdo_overlap_sets_batch: needs an explicit matrix (pcgdo_memo_create_metric(..., PCGDO_METRIC_EXPLICIT))";
        return PCGDO_ERR_UNSUPPORTED;
    }
    ctx->timed = false;
    ctx->last_launches = 0;
    if (n == 0) return PCGDO_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CK(cudaEventRecord(ctx->ev0, st));
    CK(launch_overlap_sets((const uint32_t *)memo->d_sigma, (int)memo->a, nsets, n, e1, e2, nsets == 3 ? e3 : nullptr, median,
                           cost, ctx->sm_count, st));
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    ctx->last_launches = 1;
    return PCGDO_OK;
}

extern "C" int pcgdo_overlap_sets_batch_host(pcgdo_ctx *ctx, const pcgdo_memo *memo, int nsets, size_t n, const uint32_t *e1,
                                             const uint32_t *e2, const uint32_t *e3, uint32_t *median, uint32_t *cost)
{
    if (!ctx || !memo || (nsets != 2 && nsets != 3) || (n && (!e1 || !e2 || (nsets == 3 && !e3)))) return PCGDO_ERR_ARG;
    if (n == 0) return PCGDO_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t W = (memo->a + 31) / 32, words = n * W;
    int rc;
    if ((rc = pcgdo_ensure(ctx, ctx->d_elems, 3 * words * 4)) || (rc = pcgdo_ensure(ctx, ctx->d_og, (words + n) * 4))) return rc;
    cudaStream_t st = ctx->stream;
    uint32_t *d1 = (uint32_t *)ctx->d_elems.p, *d2 = d1 + words, *d3 = d2 + words, *dm = (uint32_t *)ctx->d_og.p, *dc = dm + words;
    CK(cudaMemcpyAsync(d1, e1, words * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d2, e2, words * 4, cudaMemcpyHostToDevice, st));
    if (nsets == 3) CK(cudaMemcpyAsync(d3, e3, words * 4, cudaMemcpyHostToDevice, st));
    if ((rc = pcgdo_overlap_sets_batch_device(ctx, memo, nsets, n, d1, d2, d3, dm, dc, st))) return rc;
    if (median) CK(cudaMemcpyAsync(median, dm, words * 4, cudaMemcpyDeviceToHost, st));
    if (cost) CK(cudaMemcpyAsync(cost, dc, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PCGDO_OK;
}

namespace {
[[noreturn]] void die(const char *what, const char *detail)
{
    fprintf(stderr, "pcg_do_b200: %s: %s (no CPU fallback)\n", what, detail ? detail : "");
    abort();
}

// One context per calling thread, sized for single pairs.  Never destroyed: the threads are the host runtime's
// capabilities and live as long as the process (tearing CUDA objects down from thread-exit handlers is not safe).
pcgdo_ctx *shim_ctx(const char *who)
{
    static thread_local pcgdo_ctx *tl = nullptr;
    if (!tl) {
        if (pcgdo_create(&tl, 0) != PCGDO_OK) die(who, "cannot create a GPU context");
        tl->pairs_per_warp = 4;          // (room in the warp arena for one pair with a band of 32 diagonals per lane)
        tl->warps_per_cta = 1;
        tl->max_b = 32;
    }
    return tl;
}

// Device TCMs of the caller's cost matrices, keyed by the address of their cost table and VERIFIED on every hit
// against the symbol-change costs they were built from (the dense table is a function of those a x a entries): a
// matrix freed and re-allocated at the same address is rebuilt, not silently reused.  freeCostMtx drops the entry.
struct ShimTcm {
    pcgdo_tcm *t = nullptr;
    size_t a = 0;
    unsigned gap_open = 0;
    int variant = 0;
    std::vector<unsigned> sigma;
};
std::mutex g_mu;
std::unordered_map<const void *, ShimTcm> g_tcms[2];     // [0] linear, [1] affine

std::vector<unsigned> sigma_of(const cost_matrices_2d_t *cm)
{
    const size_t a = cm->alphSize;
    std::vector<unsigned> s(a * a);
    for (size_t i = 0; i < a; ++i)
        for (size_t j = 0; j < a; ++j) s[i * a + j] = cm->cost[(((size_t)1 << i) << a) + ((size_t)1 << j)];
    return s;
}

pcgdo_tcm *shim_tcm(pcgdo_ctx *ctx, const cost_matrices_2d_t *cm, bool affine, int variant, const char *who)
{
    std::vector<unsigned> sig = sigma_of(cm);
    std::lock_guard<std::mutex> lock(g_mu);
    ShimTcm &e = g_tcms[affine ? 1 : 0][cm->cost];
    const unsigned go = affine ? cm->gap_open_cost : 0u;
    if (e.t && (e.a != cm->alphSize || e.gap_open != go || e.variant != variant || e.sigma != sig)) {
        pcgdo_tcm_destroy(ctx, e.t);
        e.t = nullptr;
    }
    if (!e.t) {
        if (tcm_from_tables(ctx, cm->alphSize, go, cm->cost, cm->median, &e.t, variant) != PCGDO_OK) die(who, ctx->err.c_str());
        e.a = cm->alphSize; e.gap_open = go; e.variant = variant; e.sigma = std::move(sig);
    }
    return e.t;
}

// One pair through the host entry point; the three output strings come back right-aligned like dynCharToAlignIO
// leaves them (c_alignment_interface.c:544-585).
int shim_align(const char *who, bool affine, alignIO_t *in1, alignIO_t *in2, alignIO_t *gapped, alignIO_t *ungapped,
               cost_matrices_2d_t *cm, bool want)
{
    pcgdo_ctx *ctx = shim_ctx(who);
    const int variant = affine ? ctx->knobs.affine_variant : 0;
    pcgdo_tcm *t = shim_tcm(ctx, cm, affine, variant, who);
    const size_t n1 = in1->length, n2 = in2->length;
    const size_t cap = (n1 + n2 + 3) & ~(size_t)3;
    std::vector<uint8_t> elems(n1 + n2 + 1), og(cap + 4), o1(cap + 4), o2(cap + 4);
    const elem_t *s1 = in1->character + (in1->capacity - n1), *s2 = in2->character + (in2->capacity - n2);
    for (size_t i = 0; i < n1; ++i) elems[i] = (uint8_t)s1[i];
    for (size_t i = 0; i < n2; ++i) elems[n1 + i] = (uint8_t)s2[i];
    uint64_t off1 = 0, off2 = n1, out_off = 0;
    uint32_t len1 = (uint32_t)n1, len2 = (uint32_t)n2, out_len = 0;
    int32_t cost = 0;
    pcgdo_batch b{};
    b.n_pairs = 1;
    b.elems = elems.data(); b.elems_bytes = elems.size();
    b.off1 = &off1; b.off2 = &off2; b.len1 = &len1; b.len2 = &len2;
    b.out_off = &out_off; b.out_bytes = cap;
    b.cost = &cost;
    if (want) { b.out_gapped = og.data(); b.out_slot1 = o1.data(); b.out_slot2 = o2.data(); b.out_len = &out_len; }
    const size_t need_len = n1 > n2 ? n1 : n2;
    if ((int)need_len > ctx->max_len) ctx->max_len = (int)need_len;          // this thread's context only grows
    const int rc = affine ? pcgdo_align2d_affine_batch_host(ctx, t, &b) : pcgdo_align2d_batch_host(ctx, t, &b);
    if (rc != PCGDO_OK) die(who, ctx->err.c_str());
    if (want) {
        auto put = [&](alignIO_t *io, const std::vector<uint8_t> &v) {
            io->character = (elem_t *)realloc(io->character, io->capacity * sizeof(elem_t));
            if (!io->character || out_len > io->capacity) die(who, "output buffer too small");
            io->length = out_len;
            elem_t *dst = io->character + (io->capacity - out_len);
            for (uint32_t k = 0; k < out_len; ++k) dst[k] = v[k];
        };
        put(gapped, og);
        put(in1, o1);
        put(in2, o2);
        if (affine && ungapped) ungapped->length = 0;
    }
    return cost;
}
}  // namespace

extern "C" int align2d(alignIO_t *in1, alignIO_t *in2, alignIO_t *gapped, alignIO_t *ungapped,
                       cost_matrices_2d_t *cm, int getUngapped, int getGapped, int getUnion)
{
    if (getUngapped || getUnion) die("align2d", "only the (0,1,0) / (0,0,0) flag combinations PCG uses are implemented");
    if (cm->cost_model_type != 0) die("align2d", "affine cost matrix passed to the linear entry point");
    return shim_align("align2d", false, in1, in2, gapped, ungapped, cm, getGapped != 0);
}

// Reference-compatible affine entry point (EXT/c_alignment_interface.h:45-52; bound at DO/Pairwise/FFI.hsc:135-143).
// On return inputChar1/2 hold their own strings with gaps as gap_char, gappedOutput the full-length median
// (what the reference leaves there), ungappedOutput is left empty (the Haskell side never reads it, FFI.hsc:328-330).
// The environment variable PCGDO_AFFINE_VARIANT=1 selects the patched substitution lookup (see header).
extern "C" int align2dAffine(alignIO_t *in1, alignIO_t *in2, alignIO_t *gapped, alignIO_t *ungapped,
                             cost_matrices_2d_t *cm, int getMedians)
{
    if (cm->cost_model_type == 0 || cm->gap_open_cost == 0) die("align2dAffine", "linear cost matrix passed to the affine entry point");
    return shim_align("align2dAffine", true, in1, in2, gapped, ungapped, cm, getMedians != 0);
}

// EXT/c_code_alloc_setup.c:52-66: frees the tables and the struct itself; here it also drops the device copy.
extern "C" void freeCostMtx(void *input, int is_2d)
{
    if (!input) return;
    if (is_2d) {
        cost_matrices_2d_t *m = (cost_matrices_2d_t *)input;
        {
            std::lock_guard<std::mutex> lock(g_mu);
            for (auto &cache : g_tcms) {
                auto it = cache.find(m->cost);
                if (it != cache.end()) {
                    if (it->second.t) { cudaDeviceSynchronize(); pcgdo_tcm_destroy(nullptr, it->second.t); }
                    cache.erase(it);
                }
            }
        }
        free(m->cost); free(m->median); free(m->worst); free(m->prepend_cost); free(m->tail_cost);
    } else {
        cost_matrices_3d_t *m = (cost_matrices_3d_t *)input;
        free(m->cost); free(m->median);
    }
    free(input);
}

// Reference-compatible tcm-memo entry points (lib/tcm-memo/ffi/memoized-tcm/costMatrixWrapper.h:12-31, bound at
// lib/tcm-memo/src/Data/TCM/Memoized/FFI.hsc:49-69).  One lookup per call through the GPU: correct, never fast —
// the batched kernels above are the product; these keep the existing binding linkable against this library.
// Alphabets up to PCGDO_MAX_ALPHABET (the reference's script tests use 82 / 256 / 433 symbols here); elements are
// ceil(a / 64) uint64 words on the C side (dynamicCharacterOperations.h), W = ceil(a / 32) 32-bit words on the device.
extern "C" costMatrix_p matrixInit(size_t alphSize, unsigned int *tcm)
{
    pcgdo_ctx *ctx = shim_ctx("matrixInit");
    pcgdo_memo *m = nullptr;
    if (pcgdo_memo_create_metric(ctx, tcm, alphSize, PCGDO_METRIC_EXPLICIT, &m) != PCGDO_OK) die("matrixInit", ctx->err.c_str());
    return (costMatrix_p)m;
}

extern "C" void matrixDestroy(costMatrix_p untyped_ptr) { pcgdo_memo_destroy(shim_ctx("matrixDestroy"), (pcgdo_memo *)untyped_ptr); }

static unsigned int shim_lookup(const char *who, int nsets, dcElement_t *const *in, dcElement_t *retElem, costMatrix_p tcm)
{
    pcgdo_ctx *ctx = shim_ctx(who);
    const pcgdo_memo *m = (const pcgdo_memo *)tcm;
    const size_t W = (m->a + 31) / 32, W64 = (m->a + 63) / 64;
    std::vector<uint32_t> e((size_t)3 * W, 0u), med(W, 0u);
    for (int x = 0; x < nsets; ++x)
        for (size_t w = 0; w < W; ++w) e[x * W + w] = (uint32_t)(in[x]->element[w / 2] >> (32 * (w & 1)));
    uint32_t cost = 0;
    if (pcgdo_overlap_sets_batch_host(ctx, m, nsets, 1, e.data(), e.data() + W, e.data() + 2 * W, retElem ? med.data() : nullptr,
                                      &cost) != PCGDO_OK)
        die(who, ctx->err.c_str());
    if (retElem) {      // callee frees and re-allocates the median buffer (costMatrix_2d.cpp:156-166, costMatrix_3d.cpp:142-145)
        if (retElem->element != NULL) free(retElem->element);
        retElem->element = (uint64_t *)calloc(W64, sizeof(uint64_t));
        if (!retElem->element) die(who, "out of memory");
        for (size_t w = 0; w < W; ++w) retElem->element[w / 2] |= (uint64_t)med[w] << (32 * (w & 1));
    }
    return cost;
}

extern "C" unsigned int getCostAndMedian2D(dcElement_t *elem1, dcElement_t *elem2, dcElement_t *retElem, costMatrix_p tcm)
{
    dcElement_t *in[3] = {elem1, elem2, nullptr};
    return shim_lookup("getCostAndMedian2D", 2, in, retElem, tcm);
}

extern "C" unsigned int getCostAndMedian3D(dcElement_t *elem1, dcElement_t *elem2, dcElement_t *elem3, dcElement_t *retElem,
                                           costMatrix_p tcm)
{
    dcElement_t *in[3] = {elem1, elem2, elem3};
    return shim_lookup("getCostAndMedian3D", 3, in, retElem, tcm);
}
