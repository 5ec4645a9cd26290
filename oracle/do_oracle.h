/* do_oracle.h — CPU restatement of PCG's pairwise direct-optimization (DO) hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pcg_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU arm.
 *
 * Parity status:
 *   C-linear dialect  : PINNED  — checked against the reference's golden file
 *                       test/data-sets/golden-tests/chel-golden-test/chel_xml.golden
 *                       (tests/golden/chel_*.json) and fuzzed against the compiled,
 *                       unmodified reference (oracle/_ref/libpcgdo_ref.so).
 *   C-affine dialect  : parity unpinned (reference has no golden vector; checked only
 *                       against oracle/_ref, whose affine code has undefined behaviour).
 *
 * All reference citations are relative to /root/reference/.
 *   EXT = lib/core/ffi/external-direct-optimization
 *   DO  = lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization
 *   DYN = lib/core/data-structures/src/Bio/Character/Encodable/Dynamic
 */
#ifndef PCG_DO_ORACLE_H
#define PCG_DO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Dense 2^a x 2^a cost / median tables (EXT/c_code_alloc_setup.c:110-205). */
typedef struct oracle_tcm {
    int       a;        /* alphabet size including the gap symbol          */
    uint32_t  dim;      /* 1 << a                                          */
    uint32_t  gap;      /* 1 << (a-1)                                      */
    uint32_t  gap_open; /* 0 = linear                                      */
    uint32_t *cost;     /* cost  [(e1 << a) + e2]                          */
    uint32_t *median;   /* median[(e1 << a) + e2]                          */
} oracle_tcm_t;

oracle_tcm_t *oracle_tcm_new(const uint32_t *sigma /* a*a row-major */, int a, uint32_t gap_open);
void          oracle_tcm_free(oracle_tcm_t *t);

/* Band geometry of align2d (EXT/c_alignment_interface.c:101-109, EXT/alignCharacters.c:1163-1374).
 * lg / sh INCLUDE the prepended gap.  mode: 0 full, 2 = case 2, 3 = case 3 subset 3.
 * lo/hi (each lg entries) receive the inclusive column interval of every row;
 * right_ukk[i] = 1 when row i's last cell has no DELETE option. Returns #cells filled. */
uint64_t oracle_band_rows(size_t lg, size_t sh, int *mode, int32_t *lo, int32_t *hi, uint8_t *right_ukk);

/* Raw align2d (EXT/c_alignment_interface.c:16-177) with flags (0,1,0).
 * in1 / in2: element strings WITHOUT the leading gap.  Outputs (natural order, length *out_len):
 *   out_gapped : gapped medians
 *   out_slot1  : what the reference leaves in inputChar1's buffer
 *   out_slot2  : what the reference leaves in inputChar2's buffer
 *     (the slot of the LONGER input receives the aligned SHORTER string, 0 = gap)
 * cells (optional) receives the number of DP cells the reference algorithm fills. */
int64_t oracle_align2d(const oracle_tcm_t *t,
                       const uint32_t *in1, size_t n1,
                       const uint32_t *in2, size_t n2,
                       uint32_t *out_gapped, uint32_t *out_slot1, uint32_t *out_slot2,
                       size_t *out_len, uint64_t *cells);

/* Haskell wrapper common to every dialect:
 * DO/Pairwise/FFI.hsc:222-260 (algn2d), DO/Pairwise/Internal.hs:307-381,
 * DYN/Internal.hs:135-241 (deleteGaps / insertGaps), DYN/Element.hs:127-153.
 * A dynamic character is three parallel arrays (median,left,right) + a missing flag.
 * Output arrays need capacity n1+n2. Returns the cost. */
int64_t oracle_foreign_pairwise_do(const oracle_tcm_t *t,
                                   const uint32_t *m1, const uint32_t *l1, const uint32_t *r1, size_t n1, int missing1,
                                   const uint32_t *m2, const uint32_t *l2, const uint32_t *r2, size_t n2, int missing2,
                                   uint32_t *om, uint32_t *ol, uint32_t *orr, size_t *on, int *omissing,
                                   uint64_t *cells);


/* C-affine dialect: raw align2dAffine (EXT/c_alignment_interface.c:180-348, EXT/alignCharacters.c:2470-3010,
 * 3124-3256, 3483-3727) with getMedians = 1.  PARITY UNPINNED: the reference has no golden vector for it and its
 * substitution lookup has undefined behaviour (`no_gap << costMatrixDimension`, alignCharacters.c:3636).
 *   variant 0 = "as coded" on x86-64 (shift count taken mod 32: the lookup reads cost_flat[x' + y']),
 *   variant 1 = "patched"  (`<< alphSize`, the evidently intended cost[x'][y']).
 * Outputs in natural order, length *out_len: out_gapped = what lands in gappedOutput_aio (the full-length median,
 * the two median buffers are crossed at c_alignment_interface.c:314-325), out_slot1 / out_slot2 = inputChar1 /
 * inputChar2 after the call (natural order: each slot keeps ITS OWN string, gaps as gap_char).
 * cells = DP cells inside the reference's band. */
int64_t oracle_align2d_affine(const oracle_tcm_t *t, int variant,
                              const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                              uint32_t *out_gapped, uint32_t *out_slot1, uint32_t *out_slot2,
                              size_t *out_len, uint64_t *cells);

/* Haskell dialects (PARITY UNPINNED, see do_oracle.c): dialect 0 = S1 full matrix (naiveDO, naiveDOMemo,
 * unboxedSwappingDO), 1 = S1 Ukkonen ribbon (ukkonenDO), 2 = S2 Ukkonen band (unboxedUkkonenFullSpaceDO, the
 * production path for alphabets > 8).  metric_kind 0 = discrete metric, 1 = explicit sigma[a*a] through overlap.
 * in1 / in2 are median strings; outputs: median, aligned in1, aligned in2 (0 = gap). */
int64_t oracle_haskell_do(int dialect, int metric_kind, const uint32_t *sigma, int a,
                          const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                          uint32_t *out_med, uint32_t *out_a1, uint32_t *out_a2, size_t *out_len,
                          uint64_t *cells, int *final_offset);

/* Re-rooting pass for one tree and one dynamic character (reroot_oracle.c; trees only).  2*n_taxa - 3 edges are
 * reported, edge 0 = the original root edge; returns the minimum total over the edges. */
int64_t oracle_reroot(const oracle_tcm_t *t, uint32_t n_taxa, const int32_t *children,
                      const uint32_t *const *leaf, const size_t *leaf_len,
                      int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost, int64_t *edge_local, uint64_t *cells);

/* with `trial` != NULL also edge_trial[e] = Wagner edge trial: cost increase of attaching that new leaf to edge e
 * (app/pcg/PCG/Command/Build/Evaluate.hs:494-527) */
int64_t oracle_reroot_trial(const oracle_tcm_t *t, uint32_t n_taxa, const int32_t *children,
                            const uint32_t *const *leaf, const size_t *leaf_len,
                            int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost, int64_t *edge_local, uint64_t *cells,
                            const uint32_t *trial, size_t trial_len, int64_t *edge_trial);

/* Pre-order pass: implied alignments (preorder_oracle.c; PARITY UNPINNED, see its header).
 * oracle_derive_implied_alignment = deriveImpliedAlignment (DO/Internal.hs:236-286): the child's final alignment from the
 * parent's final alignment (pam/pal/par, n elements), the parent's context (left / right parts, np) and the child's
 * context (nc); 0, or -1 where the reference raises "Impossible happened".
 * oracle_preorder_tree: post-order contexts and final alignments of every node of one rooted tree (see the .c file). */
int  oracle_derive_implied_alignment(int is_left_child, uint32_t gap,
                                     const uint32_t *pam, const uint32_t *pal, const uint32_t *par, size_t n,
                                     const uint32_t *pcl, const uint32_t *pcr, size_t np,
                                     const uint32_t *ccm, const uint32_t *ccl, const uint32_t *ccr, size_t nc,
                                     uint32_t *om, uint32_t *ol, uint32_t *orr);
void oracle_lexically_disambiguate(const uint32_t *med, size_t n, int a, uint32_t *out);
int  oracle_preorder_tree(const oracle_tcm_t *t, uint32_t n_taxa, const int32_t *children,
                          const uint32_t *const *leaf, const size_t *leaf_len, size_t cap,
                          uint32_t *ctx_m, uint32_t *ctx_l, uint32_t *ctx_r, uint32_t *ctx_len,
                          uint32_t *ia_m, uint32_t *ia_l, uint32_t *ia_r, uint32_t *ia_len, uint32_t *single);

/* Alphabets of more than 32 symbols (hs_wide.c): an element is W = ceil(a/32) consecutive words, lengths count elements. */
int64_t oracle_haskell_do_w(int dialect, int metric_kind, const uint32_t *sigma, int a,
                            const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                            uint32_t *out_med, uint32_t *out_a1, uint32_t *out_a2, size_t *out_len,
                            uint64_t *cells, int *final_offset);
uint64_t oracle_overlap_w(int metric_kind, const uint32_t *sigma, int a, const uint32_t *x, const uint32_t *y, uint32_t *med);

/* re-rooting with every alignment by a Haskell dialect (alphabets > 8; elements of ceil(a/32) words) */
int64_t oracle_reroot_haskell(int dialect, int metric_kind, const uint32_t *sigma, int a, uint32_t n_taxa,
                              const int32_t *children, const uint32_t *const *leaf, const size_t *leaf_len,
                              int32_t *edge_u, int32_t *edge_v, int64_t *edge_cost, int64_t *edge_local);

/* the same with every overlap answered by `lookup` (returns the cost, writes the median) */
typedef uint64_t (*oracle_lookup_fn)(void *ctx, uint32_t x, uint32_t y, uint32_t *med);
int64_t oracle_haskell_do_lookup(int dialect, oracle_lookup_fn lookup, void *lookup_ctx, int a,
                                 const uint32_t *in1, size_t n1, const uint32_t *in2, size_t n2,
                                 uint32_t *out_med, uint32_t *out_a1, uint32_t *out_a2, size_t *out_len,
                                 uint64_t *cells, int *final_offset);

/* One overlap lookup: Sankoff median / cost of two ambiguity sets (metric_kind 0 discrete, 1 explicit sigma[a*a],
 * sigma[s*a + j] = median symbol s vs input symbol j).  Explicit branch pinned against the compiled tcm-memo. */
uint64_t oracle_overlap(int metric_kind, const uint32_t *sigma, int a, uint32_t x, uint32_t y, uint32_t *med);

#ifdef __cplusplus
}
#endif
#endif
