// do_explicit.h — internal: launch parameters of the explicit-matrix large-alphabet kernel (do_explicit.cu).
#pragma once
#include "do_metric.cuh"

namespace pcgdo {

// Batch dictionary (alphabets <= 32): when the WHOLE batch draws its elements from at most kDictMax distinct ambiguity
// sets — amino-acid strings with few ambiguity codes do — one table over those sets replaces the per-pair memo tables:
// it is staged once per CTA, 32-way replicated with copy k wholly in shared-memory bank k (the layout of the dense
// kernel, do_linear.cu), so that the 32 lanes' unrelated lookups never conflict.  The per-pair tables are 1-2 KB, too
// large to replicate per warp, and their bank conflicts (3-4 wavefronts per lookup) were the first limiter of config 4.
constexpr int kDictMax = 30;
struct DictDev {
    uint32_t count;            // distinct elements seen by dict_scan_kernel (saturates near kDictMax)
    uint32_t overflow;         // more than kDictMax of them, or a CTA's local cache overflowed
    uint32_t ok;               // set by dict_build_kernel: the table below is usable for this batch
    int32_t  emin, emax;       // range of the finite table entries (packed-fill window)
    uint32_t has_gap_median;   // some overlap of two elements has exactly the gap as its median (S2 relabels those cells)
    uint32_t pad_[2];
    uint32_t keys[64];         // open-addressing hash of the elements
    uint32_t vals[64];         // table index of keys[slot]: 2 + rank (0 = padding entry, 1 = the gap element)
    uint32_t el[32];           // index -> element
    uint32_t cL[32], gS[32];   // indel costs folded out of the cell values: overlap(el, gap), overlap(gap, el)
    int16_t  tab[32 * 32];     // [dl * 32 + ds]: 4 * (cost(dl, ds) - cL[dl] - gS[ds]) - 1, kBigSub for padding
};
constexpr uint32_t kDictSmemBytes = 65536 + 1024;     // replicated table + hash / indel arrays

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
    const DictDev      *dict;       // non-NULL: batch-dictionary mode (shared memory holds kDictSmemBytes more)
};

size_t explicit_scratch_words(int max_len, int a, size_t table_cap_entries);
cudaError_t launch_explicit(const ExplicitParams &p, int grid, int block, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_fullmatrix(const ExplicitParams &p, int grid, cudaStream_t stream);
cudaError_t launch_overlap2d(const uint32_t *sigma, int a, size_t n, const uint32_t *e1, const uint32_t *e2,
                             uint32_t *med, uint32_t *cost, int sm_count, cudaStream_t st);

cudaError_t launch_dict(const ExplicitParams &p, DictDev *dict, int grid, cudaStream_t stream);
cudaError_t launch_overlap_sets(const uint32_t *sigma, int a, int nsets, size_t n, const uint32_t *e1, const uint32_t *e2,
                                const uint32_t *e3, uint32_t *med, uint32_t *cost, int sm_count, cudaStream_t st);

}  // namespace pcgdo
