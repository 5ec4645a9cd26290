// do_multi.cu — several GPUs driven from ONE process (north_star item 4 as a library feature).
//
// PCG is a single process whose post-order sweep evaluates every dynamic character independently
// (Bio/Sequence/Block/Character.hs:440-463 hexZipMeta) and sums the characters' costs per tree
// (Bio/Sequence/Block.hs:87-106 blockCost): the (tree, character) grid shards with no exchange during the sweep.
// Here a forest is split by CHARACTER BLOCKS over the devices of a pcgdo_multi — each device keeps all node strings
// of its characters resident and runs every level locally, driven by its own host thread — and the only cross-device
// step is the sum of the per-tree partial costs, int64[n_trees] (8 KB for 1024 trees), done one of three ways:
//   HOST  one small device-to-host copy per device, added on the host;
//   PEER  one kernel on the first device that reads the other devices' sums through NVLink peer mappings;
//   NCCL  ncclAllReduce over a single-process communicator (libnccl.so.2 resolved at run time).
// Pair batches split evenly over the devices, one pipelined host call per device (no collective).
#include "do_ctx.h"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

namespace {

// the handful of NCCL entry points used, resolved with dlsym (the library must load where NCCL is absent)
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        CommInitAll = (int (*)(void **, int, const int *))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
        AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(lib, "ncclAllReduce");
        GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
        GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
        GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
    }
};
constexpr int kNcclInt64 = 4, kNcclSum = 0;      // ncclDataType_t / ncclRedOp_t values of nccl.h

__global__ void peer_sum_kernel(const int64_t *const *parts, int n_parts, uint32_t n, int64_t *out)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t s = 0;
    for (int i = 0; i < n_parts; ++i) s += parts[i][t];      // parts[i] for i > 0 are peer mappings: loads cross NVLink
    out[t] = s;
}

}  // namespace

struct pcgdo_multi {
    std::vector<int> devs;
    std::vector<pcgdo_ctx *> ctx;
    std::string err;
    bool peer_ok = false, peer_tried = false;
    Nccl nccl;
    std::vector<void *> comms;
};

struct pcgdo_mtcm {
    std::vector<pcgdo_tcm *> t;
};

struct pcgdo_mforest {
    uint32_t n_trees = 0, n_taxa = 0, n_chars = 0;
    bool by_trees = false;
    std::vector<uint32_t> lo, hi;               // character (or tree) block of each device
    std::vector<pcgdo_forest *> f;              // nullptr where a device's block is empty
    int64_t *h_stage = nullptr;                 // pinned [n_dev][n_trees]
    void *d_ptrs = nullptr, *d_sum = nullptr;   // device 0: table of the devices' cost vectors, their sum
};

#define MFAIL(m, rc_, msg)                        \
    do {                                          \
        (m)->err = (msg);                         \
        return (rc_);                             \
    } while (0)

extern "C" int pcgdo_multi_create(pcgdo_multi **out, const int *devices, int n_devices)
{
    if (!out || !devices || n_devices < 1) return PCGDO_ERR_ARG;
    *out = nullptr;
    pcgdo_multi *m = new pcgdo_multi();
    for (int i = 0; i < n_devices; ++i) {
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) { delete m; return PCGDO_ERR_ARG; }
        pcgdo_ctx *c = nullptr;
        const int rc = pcgdo_create(&c, devices[i]);
        if (rc != PCGDO_OK) {
            for (pcgdo_ctx *x : m->ctx) pcgdo_destroy(x);
            delete m;
            return rc;
        }
        m->devs.push_back(devices[i]);
        m->ctx.push_back(c);
    }
    *out = m;
    return PCGDO_OK;
}

extern "C" void pcgdo_multi_destroy(pcgdo_multi *m)
{
    if (!m) return;
    for (void *c : m->comms)
        if (c) m->nccl.CommDestroy(c);
    for (pcgdo_ctx *x : m->ctx) pcgdo_destroy(x);
    delete m;
}

extern "C" int pcgdo_multi_devices(const pcgdo_multi *m) { return m ? (int)m->ctx.size() : 0; }
extern "C" pcgdo_ctx *pcgdo_multi_ctx(pcgdo_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
extern "C" const char *pcgdo_multi_last_error(const pcgdo_multi *m) { return m ? m->err.c_str() : "null"; }

extern "C" int pcgdo_multi_configure(pcgdo_multi *m, int pairs_per_warp, int warps_per_cta, int max_len, int max_b)
{
    if (!m) return PCGDO_ERR_ARG;
    for (pcgdo_ctx *c : m->ctx) {
        int rc = pcgdo_configure(c, pairs_per_warp, warps_per_cta, max_len);
        if (rc == PCGDO_OK) rc = pcgdo_configure_ex(c, max_b, 0);
        if (rc != PCGDO_OK) MFAIL(m, rc, pcgdo_last_error(c));
    }
    return PCGDO_OK;
}

extern "C" int pcgdo_multi_tcm_create(pcgdo_multi *m, const unsigned int *sigma, size_t a, pcgdo_mtcm **out)
{
    if (!m || !sigma || !out) return PCGDO_ERR_ARG;
    pcgdo_mtcm *t = new pcgdo_mtcm();
    for (pcgdo_ctx *c : m->ctx) {
        pcgdo_tcm *x = nullptr;
        const int rc = pcgdo_tcm_create(c, sigma, a, 0, &x);
        if (rc != PCGDO_OK) {
            m->err = pcgdo_last_error(c);
            for (size_t i = 0; i < t->t.size(); ++i) pcgdo_tcm_destroy(m->ctx[i], t->t[i]);
            delete t;
            return rc;
        }
        t->t.push_back(x);
    }
    *out = t;
    return PCGDO_OK;
}

extern "C" void pcgdo_multi_tcm_destroy(pcgdo_multi *m, pcgdo_mtcm *t)
{
    if (!m || !t) return;
    for (size_t i = 0; i < t->t.size(); ++i) pcgdo_tcm_destroy(m->ctx[i], t->t[i]);
    delete t;
}

static void block_of(uint32_t n, int i, int parts, uint32_t &lo, uint32_t &hi)
{
    const uint32_t q = n / (uint32_t)parts, r = n % (uint32_t)parts;
    lo = (uint32_t)i * q + std::min<uint32_t>((uint32_t)i, r);
    hi = lo + q + ((uint32_t)i < r ? 1u : 0u);
}

extern "C" void pcgdo_mforest_destroy(pcgdo_multi *m, pcgdo_mforest *F)
{
    if (!m || !F) return;
    for (size_t i = 0; i < F->f.size(); ++i)
        if (F->f[i]) pcgdo_forest_destroy(m->ctx[i], F->f[i]);
    cudaSetDevice(m->devs[0]);
    if (F->h_stage) cudaFreeHost(F->h_stage);
    if (F->d_ptrs) cudaFree(F->d_ptrs);
    if (F->d_sum) cudaFree(F->d_sum);
    delete F;
}

extern "C" int pcgdo_mforest_create(pcgdo_multi *m, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars, uint32_t node_cap,
                                    const int32_t *children, pcgdo_mforest **out)
{
    if (!m || !children || !out || n_trees == 0 || n_taxa < 2 || n_chars == 0) return PCGDO_ERR_ARG;
    const int nd = (int)m->ctx.size();
    pcgdo_mforest *F = new pcgdo_mforest();
    F->n_trees = n_trees; F->n_taxa = n_taxa; F->n_chars = n_chars;
    F->by_trees = n_chars < (uint32_t)nd;            // fewer characters than devices: split the candidate trees instead
    F->lo.resize(nd); F->hi.resize(nd); F->f.assign(nd, nullptr);
    std::vector<int> rcs(nd, PCGDO_OK);
    std::vector<std::thread> th;
    for (int i = 0; i < nd; ++i) {
        block_of(F->by_trees ? n_trees : n_chars, i, nd, F->lo[i], F->hi[i]);
        if (F->hi[i] == F->lo[i]) continue;
        th.emplace_back([&, i] {
            if (F->by_trees)
                rcs[i] = pcgdo_forest_create(m->ctx[i], F->hi[i] - F->lo[i], n_taxa, n_chars, node_cap,
                                             children + (size_t)F->lo[i] * (n_taxa - 1) * 2, &F->f[i]);
            else
                rcs[i] = pcgdo_forest_create(m->ctx[i], n_trees, n_taxa, F->hi[i] - F->lo[i], node_cap, children, &F->f[i]);
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < nd; ++i)
        if (rcs[i] != PCGDO_OK) {
            m->err = pcgdo_last_error(m->ctx[i]);
            const int rc = rcs[i];
            pcgdo_mforest_destroy(m, F);
            return rc;
        }
    if (cudaSetDevice(m->devs[0]) != cudaSuccess ||
        cudaHostAlloc((void **)&F->h_stage, (size_t)nd * n_trees * 8, cudaHostAllocPortable) != cudaSuccess ||
        cudaMalloc(&F->d_ptrs, (size_t)nd * sizeof(void *)) != cudaSuccess ||
        cudaMalloc(&F->d_sum, (size_t)n_trees * 8) != cudaSuccess) {
        m->err = std::string("pcgdo_mforest_create: ") + cudaGetErrorString(cudaGetLastError());
        pcgdo_mforest_destroy(m, F);
        return PCGDO_ERR_CUDA;
    }
    *out = F;
    return PCGDO_OK;
}

extern "C" int pcgdo_mforest_block(const pcgdo_mforest *F, int i, uint32_t *lo, uint32_t *hi, int *by_trees)
{
    if (!F || i < 0 || i >= (int)F->f.size()) return PCGDO_ERR_ARG;
    if (lo) *lo = F->lo[i];
    if (hi) *hi = F->hi[i];
    if (by_trees) *by_trees = F->by_trees ? 1 : 0;
    return PCGDO_OK;
}

// run fn(i) for every device with a non-empty block, one host thread each; first failure wins
template <class Fn>
static int for_each_device(pcgdo_multi *m, pcgdo_mforest *F, Fn fn)
{
    const int nd = (int)m->ctx.size();
    std::vector<int> rcs(nd, PCGDO_OK);
    std::vector<std::thread> th;
    for (int i = 1; i < nd; ++i)
        if (F->f[i]) th.emplace_back([&, i] { rcs[i] = fn(i); });
    if (F->f[0]) rcs[0] = fn(0);
    for (auto &t : th) t.join();
    for (int i = 0; i < nd; ++i)
        if (rcs[i] != PCGDO_OK) MFAIL(m, rcs[i], pcgdo_last_error(m->ctx[i]));
    return PCGDO_OK;
}

extern "C" int pcgdo_mforest_set_leaves(pcgdo_multi *m, pcgdo_mforest *F, const uint8_t *elems, const uint64_t *off,
                                        const uint32_t *len)
{
    if (!m || !F || !elems || !off || !len) return PCGDO_ERR_ARG;
    return for_each_device(m, F, [&](int i) -> int {
        if (F->by_trees) return pcgdo_forest_set_leaves(m->ctx[i], F->f[i], elems, off, len);
        const uint32_t c0 = F->lo[i], nc = F->hi[i] - F->lo[i];
        std::vector<uint64_t> o((size_t)F->n_taxa * nc);
        std::vector<uint32_t> l((size_t)F->n_taxa * nc);
        for (uint32_t v = 0; v < F->n_taxa; ++v)
            for (uint32_t c = 0; c < nc; ++c) {
                o[(size_t)v * nc + c] = off[(size_t)v * F->n_chars + c0 + c];
                l[(size_t)v * nc + c] = len[(size_t)v * F->n_chars + c0 + c];
            }
        return pcgdo_forest_set_leaves(m->ctx[i], F->f[i], elems, o.data(), l.data());
    });
}

extern "C" int pcgdo_mforest_set_trees(pcgdo_multi *m, pcgdo_mforest *F, const int32_t *children)
{
    if (!m || !F || !children) return PCGDO_ERR_ARG;
    return for_each_device(m, F, [&](int i) -> int {
        const int32_t *ch = F->by_trees ? children + (size_t)F->lo[i] * (F->n_taxa - 1) * 2 : children;
        return pcgdo_forest_set_trees(m->ctx[i], F->f[i], ch);
    });
}

extern "C" int pcgdo_mforest_enable_rerooting(pcgdo_multi *m, pcgdo_mforest *F)
{
    if (!m || !F) return PCGDO_ERR_ARG;
    return for_each_device(m, F, [&](int i) -> int { return pcgdo_forest_enable_rerooting(m->ctx[i], F->f[i], 0); });
}

static bool enable_peers(pcgdo_multi *m)
{
    if (m->peer_tried) return m->peer_ok;
    m->peer_tried = true;
    if (cudaSetDevice(m->devs[0]) != cudaSuccess) return false;
    for (size_t i = 1; i < m->devs.size(); ++i) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, m->devs[0], m->devs[i]) != cudaSuccess || !can) return false;
        const cudaError_t e = cudaDeviceEnablePeerAccess(m->devs[i], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
        cudaGetLastError();
    }
    m->peer_ok = true;
    return true;
}

extern "C" int pcgdo_mforest_score(pcgdo_multi *m, const pcgdo_mtcm *tcm, pcgdo_mforest *F, int rerooted, int reduce_mode,
                                   int64_t *tree_costs, double *wall_ms, float *sweep_ms, double *reduce_ms)
{
    if (!m || !tcm || !F || !tree_costs) return PCGDO_ERR_ARG;
    const int nd = (int)m->ctx.size();
    if (reduce_mode == PCGDO_REDUCE_PEER && nd > 1 && !F->by_trees && !enable_peers(m))
        MFAIL(m, PCGDO_ERR_UNSUPPORTED, "peer access between the devices is not available");
    if (reduce_mode == PCGDO_REDUCE_NCCL && nd > 1 && !F->by_trees && m->comms.empty()) {
        if (!m->nccl.load()) MFAIL(m, PCGDO_ERR_UNSUPPORTED, "libnccl.so.2 cannot be loaded");
        m->comms.assign(nd, nullptr);
        const int r = m->nccl.CommInitAll(m->comms.data(), nd, m->devs.data());
        if (r != 0) {
            m->comms.clear();
            MFAIL(m, PCGDO_ERR_CUDA, std::string("ncclCommInitAll: ") + (m->nccl.GetErrorString ? m->nccl.GetErrorString(r) : "failed"));
        }
    }
    const auto t0 = std::chrono::steady_clock::now();
    const bool host_mode = reduce_mode == PCGDO_REDUCE_HOST || F->by_trees || nd == 1;
    std::vector<float> ms(nd, 0.f);
    // every device sweeps its block from its own host thread (the calls return when the device is done)
    int rc = for_each_device(m, F, [&](int i) -> int {
        pcgdo_ctx *c = m->ctx[i];
        int r = rerooted ? pcgdo_forest_score_rerooted(c, tcm->t[i], F->f[i], nullptr) : pcgdo_forest_score(c, tcm->t[i], F->f[i], nullptr);
        if (r != PCGDO_OK) return r;
        pcgdo_last_kernel_ms(c, &ms[i], nullptr);
        if (host_mode) r = pcgdo_forest_tree_costs(c, F->f[i], F->h_stage + (size_t)i * F->n_trees);
        return r;
    });
    if (rc != PCGDO_OK) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    if (F->by_trees) {                                  // concatenation, no sum
        for (int i = 0; i < nd; ++i)
            if (F->f[i]) std::copy(F->h_stage + (size_t)i * F->n_trees, F->h_stage + (size_t)i * F->n_trees + (F->hi[i] - F->lo[i]),
                                   tree_costs + F->lo[i]);
    } else if (host_mode) {
        std::fill(tree_costs, tree_costs + F->n_trees, (int64_t)0);
        for (int i = 0; i < nd; ++i)
            if (F->f[i])
                for (uint32_t t = 0; t < F->n_trees; ++t) tree_costs[t] += F->h_stage[(size_t)i * F->n_trees + t];
    } else if (reduce_mode == PCGDO_REDUCE_PEER) {
        pcgdo_ctx *c0 = m->ctx[0];
        std::vector<const void *> parts;
        for (int i = 0; i < nd; ++i)
            if (F->f[i]) parts.push_back(pcgdo_forest_tree_cost_ptr(F->f[i]));
        if (cudaSetDevice(m->devs[0]) != cudaSuccess ||
            cudaMemcpyAsync(F->d_ptrs, parts.data(), parts.size() * sizeof(void *), cudaMemcpyHostToDevice, c0->stream) != cudaSuccess)
            MFAIL(m, PCGDO_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
        peer_sum_kernel<<<(F->n_trees + 255) / 256, 256, 0, c0->stream>>>((const int64_t *const *)F->d_ptrs, (int)parts.size(),
                                                                         F->n_trees, (int64_t *)F->d_sum);
        if (cudaGetLastError() != cudaSuccess ||
            cudaMemcpyAsync(tree_costs, F->d_sum, (size_t)F->n_trees * 8, cudaMemcpyDeviceToHost, c0->stream) != cudaSuccess ||
            cudaStreamSynchronize(c0->stream) != cudaSuccess)
            MFAIL(m, PCGDO_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
    } else if (reduce_mode == PCGDO_REDUCE_NCCL) {
        int r = m->nccl.GroupStart();
        for (int i = 0; i < nd && r == 0; ++i) {
            void *p = pcgdo_forest_tree_cost_ptr(F->f[i]);
            r = m->nccl.AllReduce(p, p, F->n_trees, kNcclInt64, kNcclSum, m->comms[i], m->ctx[i]->stream);
        }
        const int r2 = m->nccl.GroupEnd();
        if (r != 0 || r2 != 0) MFAIL(m, PCGDO_ERR_CUDA, "ncclAllReduce failed");
        if (cudaSetDevice(m->devs[0]) != cudaSuccess ||
            cudaMemcpyAsync(tree_costs, pcgdo_forest_tree_cost_ptr(F->f[0]), (size_t)F->n_trees * 8, cudaMemcpyDeviceToHost,
                            m->ctx[0]->stream) != cudaSuccess)
            MFAIL(m, PCGDO_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
        for (int i = 0; i < nd; ++i) {
            if (cudaSetDevice(m->devs[i]) != cudaSuccess || cudaStreamSynchronize(m->ctx[i]->stream) != cudaSuccess)
                MFAIL(m, PCGDO_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
        }
    } else {
        MFAIL(m, PCGDO_ERR_ARG, "unknown reduce_mode");
    }
    const auto t2 = std::chrono::steady_clock::now();
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(t2 - t0).count();
    if (reduce_ms) *reduce_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    if (sweep_ms) *sweep_ms = *std::max_element(ms.begin(), ms.end());
    return PCGDO_OK;
}

// ---------------------------------------------------------------------------------------------- pair batches
extern "C" int pcgdo_multi_align2d_batch_host(pcgdo_multi *m, const pcgdo_mtcm *tcm, const pcgdo_batch *b)
{
    if (!m || !tcm || !b) return PCGDO_ERR_ARG;
    if (b->n_pairs == 0) return PCGDO_OK;
    const int nd = (int)std::min<size_t>(m->ctx.size(), b->n_pairs);
    const bool want_str = b->out_gapped || b->out_slot1 || b->out_slot2;
    if (want_str && !b->out_off) MFAIL(m, PCGDO_ERR_ARG, "out_off required when aligned strings are requested");
    std::vector<int> rcs(nd, PCGDO_OK);
    auto run = [&](int i) {
        uint32_t lo, hi;
        block_of((uint32_t)b->n_pairs, i, nd, lo, hi);
        const size_t k = hi - lo;
        // this device's share as a batch of its own: string / output ranges re-based so that only its bytes travel
        uint64_t emin = UINT64_MAX, emax = 0, omin = UINT64_MAX, omax = 0;
        for (size_t p = lo; p < hi; ++p) {
            emin = std::min(emin, std::min(b->off1[p], b->off2[p]));
            emax = std::max(emax, std::max(b->off1[p] + b->len1[p], b->off2[p] + b->len2[p]));
            if (want_str) {
                omin = std::min(omin, b->out_off[p]);
                omax = std::max(omax, b->out_off[p] + (((uint64_t)b->len1[p] + b->len2[p] + 3) & ~(uint64_t)3));
            }
        }
        if (emax > b->elems_bytes || (want_str && omax > b->out_bytes)) { rcs[i] = PCGDO_ERR_ARG; return; }
        std::vector<uint64_t> o1(k), o2(k), oo(want_str ? k : 0);
        for (size_t p = 0; p < k; ++p) {
            o1[p] = b->off1[lo + p] - emin;
            o2[p] = b->off2[lo + p] - emin;
            if (want_str) oo[p] = b->out_off[lo + p] - omin;
        }
        pcgdo_batch s = *b;
        s.n_pairs = k;
        s.elems = b->elems + emin; s.elems_bytes = emax - emin;
        s.off1 = o1.data(); s.off2 = o2.data();
        s.len1 = b->len1 + lo; s.len2 = b->len2 + lo;
        s.out_off = want_str ? oo.data() : nullptr;
        s.out_bytes = want_str ? omax - omin : 0;
        s.out_gapped = b->out_gapped ? b->out_gapped + omin : nullptr;
        s.out_slot1 = b->out_slot1 ? b->out_slot1 + omin : nullptr;
        s.out_slot2 = b->out_slot2 ? b->out_slot2 + omin : nullptr;
        s.out_len = b->out_len ? b->out_len + lo : nullptr;
        s.cost = b->cost + lo;
        s.cells = b->cells ? b->cells + lo : nullptr;
        s.out_off_actual = b->out_off_actual ? b->out_off_actual + lo : nullptr;
        rcs[i] = pcgdo_align2d_batch_host(m->ctx[i], tcm->t[i], &s);
        if (rcs[i] == PCGDO_OK && s.out_off_actual)
            for (size_t p = 0; p < k; ++p) s.out_off_actual[p] += omin;
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nd; ++i) th.emplace_back(run, i);
    run(0);
    for (auto &t : th) t.join();
    for (int i = 0; i < nd; ++i)
        if (rcs[i] != PCGDO_OK) MFAIL(m, rcs[i], rcs[i] == PCGDO_ERR_ARG ? "batch layout out of bounds" : pcgdo_last_error(m->ctx[i]));
    return PCGDO_OK;
}
