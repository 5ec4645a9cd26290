// do_ctx.h — internal: context / TCM objects shared by do_api.cu and do_forest.cu.
#pragma once
#include <string>
#include <vector>

#include "do_device.cuh"
#include "do_host.h"

struct pcgdo_tcm {
    pcgdo::TcmDev dev{};
    size_t a = 0;
    unsigned int gap_open = 0;
    void *d_aff_sub = nullptr;       // affine only: u16 [s' << (a-1) | l']
    int affine_variant = 0;
    void *d_sub4 = nullptr, *d_sub4_nat = nullptr, *d_median = nullptr, *d_cgap = nullptr, *d_gapc = nullptr;
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct pcgdo_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    size_t host_chunk_pairs = 0;     // pairs per pipelined chunk of the host entry point (0 = 65536)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    bool hold_ev0 = false;           // an enclosing host call owns ev0 (its timing spans the nested device call)
    int last_launches = 0;
    // tunables
    int pairs_per_warp = 32;
    int warps_per_cta = 32;
    int max_len = 1024;
    int max_b = 8;
    size_t arena_words = 0;          // 0 = derive from max_len
    // experiment switches, read from the environment ONCE (pcgdo_create / pcgdo_configure_env), never per call
    struct Knobs {
        size_t host_chunk_pairs = 0;   // PCGDO_HOST_CHUNK_PAIRS   pairs per pipelined chunk of the host entry point
        size_t h2d_ahead = 0;          // PCGDO_H2D_AHEAD          chunks whose strings are queued ahead (0 = all)
        bool zero_copy = false;        // PCGDO_ZERO_COPY          emit straight into pinned host arrays
        bool trace_pipeline = false;   // PCGDO_TRACE_PIPELINE     per-chunk completion times to stderr
        bool no_pack16 = false;        // PCGDO_NO_PACK16          32-bit fills only
        bool no_quad = false;          // PCGDO_NO_QUAD            dense path: no four-pairs-per-warp kernel
        unsigned quad_flags = 0;       // PCGDO_QUAD_FLAGS         bit 0: stagger first batches; bit 1: phase cycle counts to stderr;
                                       //                          bit 3 / 4 / 6: without the 4- / 6- / 5-pairs-per-warp shape; bit 5: with the 8-pairs one
        bool affine_rowwise = false;   // PCGDO_AFFINE_ROWWISE     first-generation affine fill
        bool metric_registers = false; // PCGDO_METRIC_REGISTERS   do_metric.cu instead of the table kernel
        bool explicit_multi_pass = false;  // PCGDO_EXPLICIT_MULTI_PASS
        bool no_dict = false;          // PCGDO_NO_DICT            large alphabets: per-pair tables only (no batch dictionary)
        int  skip_ahead = 64;          // PCGDO_SKIP_FACTOR / PCGDO_NO_SKIP_AHEAD (0)
        unsigned skip_b = 0;           // PCGDO_SKIP_B
        int  affine_variant = 0;       // PCGDO_AFFINE_VARIANT     reading of the reference's UB shift in the align2dAffine shim
    } knobs;
    // device scratch
    DevBuf arena, counters, overflow, big, singles, singles2, phase, gen_scratch, gen_off, tscratch, d_dict;
    // staging for the host-buffer entry point
    DevBuf d_elems, d_off1, d_off2, d_len1, d_len2, d_out_off, d_og, d_o1, d_o2, d_olen, d_cost, d_cells;
    DevBuf d_strag;                  // descriptors of the few pairs a pipelined batch hands to the general-geometry kernel
    // compact transfer: ring of packed output slots, per-pair actual offsets, per-chunk byte counts
    DevBuf d_stage, d_actual, d_total;
    unsigned long long *h_total = nullptr;   // pinned
    size_t h_total_cap = 0;
    std::string err;
};

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
            return PCGDO_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)


int pcgdo_ensure(pcgdo_ctx *ctx, DevBuf &b, size_t bytes);

// Enqueue the narrow (+ wide) warp kernels for one batch on `st`.  Work counters are reset; the overflow
// counter only when `reset_overflow` (the forest scheduler accumulates it over all levels and checks once).
// *out_p receives the parameters used (for the caller's overflow handling).
int pcgdo_enqueue_linear(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo::BatchDev &bd, cudaStream_t st,
                         bool reset_overflow, pcgdo::LinearParams *out_p, bool *wide_ok_out);

// Large alphabets (do_explicit.cu): enqueue the per-pair table kernel for a batch described on the device.
struct pcgdo_memo;
namespace pcgdo { struct MetricBatchDev; }
int pcgdo_enqueue_table(pcgdo_ctx *ctx, const pcgdo_memo *memo, int alphabet, int dialect, const pcgdo::MetricBatchDev &bd,
                        cudaStream_t st, bool reset_overflow, unsigned int **counters_out, bool try_dict = false);
