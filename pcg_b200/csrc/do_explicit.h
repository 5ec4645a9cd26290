// do_explicit.h — internal: launch parameters of the explicit-matrix large-alphabet kernel (do_explicit.cu).
#pragma once
#include "do_metric.cuh"

namespace pcgdo {

struct ExplicitParams {
    MetricBatchDev b;
    int       a;              // alphabet size (incl. gap), <= 32
    int       dialect;        // 0 S1 full, 1 S1 Ukkonen, 2 S2 Ukkonen (3 = unboxedFullMatrixDO: its own kernel)
    int       kind;           // 0 discrete metric (closed form, sigma unused), 2 explicit sigma
    int       pack16;         // allow the packed 16-bit fill
    int       skip_ahead;     // dialect 2: let the doubling schedule jump levels it can prove would double, to a band of at
                              // most this many times the current one's cells (0 = plain level-by-level doubling)
    uint32_t  skip_b;         // bit mask of diagonals-per-lane values (2, 4, 8, 16) not to use: round up instead
    const uint32_t *sigma;    // device [a*a]: sigma[s*a + j], s = median symbol, j = input symbol
    uint32_t  delta;          // cheapest indel of a single non-gap symbol (Ukkonen coefficient)
    uint32_t *arena;          // per warp: direction words + op streams
    uint32_t  arena_words;
    uint32_t *tscratch;       // per warp: hash tables, distinct elements, profiles, the pair's table
    uint32_t  tscratch_words;
    uint32_t  seq_cap;        // bytes of staging per warp
    uint32_t  smem_tab_entries;   // int16 entries of shared memory per warp for small tables
    int       pairs_per_warp;
    unsigned int *counter;
    unsigned int *overflow_count;
    uint32_t  k_four, k_minus1, k_one, k_64k;
    // multi-pass mode (one Ukkonen offset per launch); all NULL = every warp walks the whole schedule of its pairs
    const uint32_t     *in_list;    // pairs of this pass (NULL in the first pass: all pairs)
    const unsigned int *in_count;
    uint32_t           *out_list;   // pairs whose band must double: the next pass
    unsigned int       *out_count;
    long long          *pair_off;   // [n_pairs] offset a queued pair continues with
};

size_t explicit_scratch_words(int max_len, int a, size_t table_cap_entries);
cudaError_t launch_explicit(const ExplicitParams &p, int grid, int block, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_fullmatrix(const ExplicitParams &p, int grid, cudaStream_t stream);
cudaError_t launch_overlap2d(const uint32_t *sigma, int a, size_t n, const uint32_t *e1, const uint32_t *e2,
                             uint32_t *med, uint32_t *cost, int sm_count, cudaStream_t st);

cudaError_t launch_overlap_sets(const uint32_t *sigma, int a, int nsets, size_t n, const uint32_t *e1, const uint32_t *e2,
                                const uint32_t *e3, uint32_t *med, uint32_t *cost, int sm_count, cudaStream_t st);

}  // namespace pcgdo
