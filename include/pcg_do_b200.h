/* pcg_do_b200.h — C ABI of the B200-native direct-optimization (DO) backend.
 *
 * Drop-in boundary for PCG's pairwise DO path: plain C types only (no torch, no CUDA types in
 * the signatures).  Two groups of entry points:
 *
 *  (1) REFERENCE-COMPATIBLE symbols — same name, argument order, struct layout and ownership
 *      rules as the C objects PCG's Haskell code binds today, so the existing
 *      `foreign import ccall` declarations keep working against this library:
 *        setUp2dCostMtx  <- EXT/c_code_alloc_setup.h / .c:110-205   (bound at Data/TCM/Dense/FFI.hsc:319-334)
 *        align2d         <- EXT/c_alignment_interface.h:29-38, .c:16-177 (bound at DO/Pairwise/FFI.hsc:122-132)
 *      (EXT = /root/reference/lib/core/ffi/external-direct-optimization,
 *       DO  = /root/reference/lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization)
 *      They run a batch of one pair through the GPU.  Correct, never fast: the FFI call, two
 *      copies and a launch per pair dominate.
 *
 *  (2) BATCHED entry points (`pcgdo_*`) — what a `foreignPairwiseDOBatch`-style binding and the
 *      post-order level scheduler call: many pairs per launch, data either in host buffers
 *      (copies inside the call) or already resident in HBM.
 *
 * There is NO CPU fallback: every entry point fails loudly (non-zero return / abort for the
 * reference-compatible `int cost` signatures) when no CUDA device is usable.
 */
#ifndef PCG_DO_B200_H
#define PCG_DO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- reference-compatible types */

typedef unsigned int elem_t;                       /* EXT/dyn_character.h:30 */

/* EXT/c_alignment_interface.h:14-18 — payload lives in the LAST `length` slots of `character`. */
typedef struct alignIO_t {
    elem_t *character;
    size_t  length;
    size_t  capacity;
} alignIO_t;

/* EXT/costMatrix.h:35-93 — field order and types must not change (hsc #peek/#poke at
 * Data/TCM/Dense/FFI.hsc:170-233 read alphSize, cost_model_type, cost, median directly). */
typedef struct cost_matrices_2d_t {
    size_t        alphSize;
    size_t        costMatrixDimension;
    elem_t        gap_char;
    int           cost_model_type;
    int           include_ambiguities;
    unsigned int  gap_open_cost;
    int           is_metric;
    size_t        num_elements;
    unsigned int *cost;
    elem_t       *median;
    unsigned int *worst;
    unsigned int *prepend_cost;
    unsigned int *tail_cost;
} cost_matrices_2d_t;

/* EXT/costMatrix.h:102-130 — what setUp3dCostMtx fills.  PCG builds it next to every 2D matrix
 * (Data/TCM/Dense/FFI.hsc:461-476) although the 3D aligner itself is disabled upstream (DO/Pairwise.hs:18). */
typedef struct cost_matrices_3d_t {
    size_t        alphSize;
    size_t        costMatrixDimension;
    elem_t        gap_char;
    int           cost_model_type;
    int           include_ambiguities;
    unsigned int  gap_open_cost;
    size_t        num_elements;
    unsigned int *cost;
    elem_t       *median;
} cost_matrices_3d_t;

/* Builds the dense 2^a x 2^a cost / median tables on the host (same loops and the same
 * sigma[pos*a + nucleotide] indexing as the reference) into caller-provided storage. */
void setUp2dCostMtx(cost_matrices_2d_t *retMtx, unsigned int *tcm, size_t alphSize, unsigned int gap_open);

/* EXT/c_code_alloc_setup.c:208-270 — the three-way tables, (2^a + 1)^3 entries each, indexed
 * (((e1 << a) + e2) << a) + e3 (EXT/costMatrix.c:111-120): setup-time host work like setUp2dCostMtx, bound at
 * Data/TCM/Dense/FFI.hsc:327-334.  Same sums of `distance` and the same median-set rule as the reference. */
void setUp3dCostMtx(cost_matrices_3d_t *retMtx, unsigned int *tcm, size_t alphSize, unsigned int gap_open);

/* EXT/c_code_alloc_setup.c:52-66: frees the tables AND the struct (both malloc'ed, as PCG allocates them); also drops
 * the device copy the align2d / align2dAffine shims cache for a 2D matrix. */
void freeCostMtx(void *input, int is_2d);

/* Pairwise DO of one pair on the GPU.  Same contract as the reference's align2d for the flag
 * combination PCG uses (getUngapped=0, getGapped=1, getUnion=0): on return
 *   inputChar1_aio / inputChar2_aio hold the two aligned strings (the slot of the LONGER input
 *   receives the aligned SHORTER string, gaps as 0), gappedOutput_aio the gapped medians,
 *   all right-aligned with `length` updated; the return value is the alignment cost.
 * getUngapped / getUnion other than 0 abort (PCG never passes them). */
int align2d(alignIO_t *inputChar1_aio, alignIO_t *inputChar2_aio,
            alignIO_t *gappedOutput_aio, alignIO_t *ungappedOutput_aio,
            cost_matrices_2d_t *costMtx2d, int getUngapped, int getGapped, int getUnion);

/* Affine twin of align2d (EXT/c_alignment_interface.h:45-52).  inputChar1/2 come back holding their own strings
 * with gaps as gap_char, gappedOutput the full-length median; ungappedOutput is returned empty (PCG's binding
 * never reads it).  See pcgdo_tcm_create_affine below for the `variant` question (env PCGDO_AFFINE_VARIANT). */
int align2dAffine(alignIO_t *inputChar1_aio, alignIO_t *inputChar2_aio,
                  alignIO_t *gappedOutput_aio, alignIO_t *ungappedOutput_aio,
                  cost_matrices_2d_t *costMtx2d_affine, int getMedians);

/* ---------------------------------------------------------------- batched API */

typedef struct pcgdo_ctx pcgdo_ctx;     /* one per (process, GPU): streams, scratch, work counters */
typedef struct pcgdo_tcm pcgdo_tcm;     /* device copy of one dense TCM                            */

enum {
    PCGDO_OK = 0,
    PCGDO_ERR_CUDA = 1,          /* no device / CUDA call failed (message in pcgdo_last_error) */
    PCGDO_ERR_ARG = 2,
    PCGDO_ERR_UNSUPPORTED = 3,   /* alphabet > 8, TCM cost too large for the 16-bit device table, ... */
    PCGDO_ERR_OVERFLOW = 4       /* a pair needs more scratch than the configured slot size   */
};

int         pcgdo_create(pcgdo_ctx **out, int device);
void        pcgdo_destroy(pcgdo_ctx *ctx);
const char *pcgdo_last_error(const pcgdo_ctx *ctx);

/* Tunables (call before the first batch; 0 keeps the default).
 *   pairs_per_warp : pairs a warp fills back-to-back before tracing them lane-parallel (1..32)
 *   warps_per_cta  : 1..32
 *   max_len        : longest string (elements, without the leading gap) the fast path must hold
 *                    in shared memory; longer pairs take the general-geometry kernel            */
int pcgdo_configure(pcgdo_ctx *ctx, int pairs_per_warp, int warps_per_cta, int max_len);
/* The experiment switches of DESIGN.md (PCGDO_* environment variables) are read once, in pcgdo_create; this re-reads
 * them (tests flip them between calls).  No entry point consults the environment on its own. */
int pcgdo_configure_env(pcgdo_ctx *ctx);

/* sigma: a*a row-major symbol-change costs, gap is symbol a-1.  a <= 8. */
int  pcgdo_tcm_create(pcgdo_ctx *ctx, const unsigned int *sigma, size_t a, unsigned int gap_open, pcgdo_tcm **out);
void pcgdo_tcm_destroy(pcgdo_ctx *ctx, pcgdo_tcm *tcm);

/* One batch of raw align2d problems.  Elements are one byte each (a <= 8).
 * Pair k: first string = elems[off1[k] .. off1[k]+len1[k]), second likewise (no leading gap).
 * Outputs for pair k start at byte out_off[k] (must be a multiple of 4, with room for
 * round_up(len1[k]+len2[k], 4) bytes) of each of out_gapped / out_slot1 / out_slot2
 * (same meaning as align2d's gappedOutput / inputChar1 / inputChar2 after the call);
 * out_len[k] = alignment length, cost[k] = alignment cost, cells[k] (may be NULL) = DP cells of
 * the reference algorithm's band for that pair.  out_slot1/out_slot2/out_gapped may be NULL
 * (cost-only).  All pointers are DEVICE pointers for *_device, HOST pointers for *_host. */
typedef struct pcgdo_batch {
    size_t          n_pairs;
    const uint8_t  *elems;
    size_t          elems_bytes;     /* host variant: bytes to copy in                      */
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    const uint64_t *out_off;
    size_t          out_bytes;       /* host variant: bytes to copy back per output array   */
    uint8_t        *out_gapped, *out_slot1, *out_slot2;
    uint32_t       *out_len;
    int32_t        *cost;
    uint64_t       *cells;
    /* HOST entry point only, may be NULL.  Non-NULL selects COMPACT TRANSFER: out_off[] still reserves each pair's
     * capacity (and must be monotone), but only the bytes the alignments produced cross the PCIe link, packed per
     * pipelined chunk inside the chunk's reserved region; pair k's three outputs are then at out_off_actual[k] of
     * out_gapped / out_slot1 / out_slot2 — not at out_off[k].  With pinned out arrays (pcgdo_host_alloc, or any
     * cudaHostAlloc'ed / registered memory) the kernel stores them there directly over PCIe, whole 128-byte lines at a time. */
    uint64_t       *out_off_actual;
} pcgdo_batch;

/* Pinned, mapped host memory for batch buffers (cudaHostAlloc): transfers to and from it run at the full PCIe rate,
 * and with out_off_actual set the kernel writes the aligned strings straight into it (zero-copy, no staging). */
void *pcgdo_host_alloc(pcgdo_ctx *ctx, size_t bytes);
void  pcgdo_host_free(pcgdo_ctx *ctx, void *p);

/* Enqueue on `cuda_stream` (a cudaStream_t passed as void*, NULL = the context's stream);
 * asynchronous unless a pair overflows the fast path's scratch. */
int pcgdo_align2d_batch_device(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b, void *cuda_stream);

/* Host buffers in, host buffers out: H2D copies, kernels and D2H copies inside the call. */
int pcgdo_align2d_batch_host(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b);

/* ---------------------------------------------------------------- affine gap costs
 * align2dAffine (EXT/c_alignment_interface.c:180-348) per pair: out_gapped = what the reference leaves in
 * gappedOutput_aio (the full-length median), out_slot1 / out_slot2 = inputChar1 / inputChar2 after the call
 * (each slot keeps its own string, gaps as gap_char).  The reference's substitution lookup is undefined
 * behaviour (alignCharacters.c:3636 shifts by 2^a); `variant` 0 reproduces the reference as compiled for x86-64
 * (shift count mod 32), 1 the evidently intended `<< alphSize`.  PCG itself never enables affine costs
 * (gap-open is hard-wired to 0, Bio/Metadata/Dynamic/Internal.hs:346); upstream has no golden vector for it. */
int pcgdo_tcm_create_affine(pcgdo_ctx *ctx, const unsigned int *sigma, size_t a, unsigned int gap_open, int variant,
                            pcgdo_tcm **out);
int pcgdo_align2d_affine_batch_device(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b, void *cuda_stream);
int pcgdo_align2d_affine_batch_host(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, const pcgdo_batch *b);

/* ---------------------------------------------------------------- large alphabets (a > 8): Haskell dialects
 * The reference dispatches alphabets of more than 8 symbols to its Haskell aligners over an overlap function
 * (DO/Internal.hs:67-83).  Elements are 32-bit ambiguity sets here (a <= 24 on this path); the overlap is the
 * DISCRETE METRIC's closed form (Data/MetricRepresentation.hs:116-128), computed per cell on the device.
 *   dialect 0 = S1 full matrix   : naiveDO / naiveDOMemo / unboxedSwappingDO
 *   dialect 1 = S1 Ukkonen ribbon: ukkonenDO
 *   dialect 2 = S2 Ukkonen band  : unboxedUkkonenFullSpaceDO  (PCG's production path for a > 8)
 *   dialect 3 = S2 full matrix   : unboxedFullMatrixDO, with its first column seeded from row 0 as coded upstream
 *                                  (batch entry points only, a <= 32; a plain kernel — it is not on PCG's dispatch path)
 * Inputs are median strings (what `measureAndUngapCharacters` leaves); outputs per pair: median, aligned first
 * input, aligned second input (0 = gap), at element offset out_off[k] (room for len1+len2 elements).
 * Upstream has neither golden vectors nor (here) an executable for these paths: parity is against the restatement
 * in oracle/do_oracle.c, itself pinned end to end by the reference's script-test DAG costs.  Explicit
 * matrices, the additive (L1) closed form and alphabets of more than 32 symbols: pcgdo_explicit_batch_* below.
 * elems_count / out_count count 32-bit WORDS (= elements up to 32 symbols). */
typedef struct pcgdo_batch32 {
    size_t          n_pairs;
    const uint32_t *elems;
    size_t          elems_count;     /* host variant: 32-bit words to copy in          */
    const uint64_t *off1, *off2;     /* element offsets                                */
    const uint32_t *len1, *len2;
    const uint64_t *out_off;
    size_t          out_count;       /* host variant: 32-bit words to copy back per array */
    uint32_t       *out_med, *out_a1, *out_a2;
    uint32_t       *out_len;
    int32_t        *cost;
    uint64_t       *cells;           /* may be NULL: cells of the final band           */
    int32_t        *final_offset;    /* may be NULL: Ukkonen offset reached (-1 = full matrix) */
} pcgdo_batch32;
int pcgdo_metric_batch_device(pcgdo_ctx *ctx, int alphabet, int dialect, const pcgdo_batch32 *b, void *cuda_stream);
int pcgdo_metric_batch_host(pcgdo_ctx *ctx, int alphabet, int dialect, const pcgdo_batch32 *b);

/* ---------------------------------------------------------------- explicit matrices for a > 8: tcm-memo replacement
 * For alphabets the dense table cannot hold and a symbol-change matrix that is neither the discrete metric nor
 * additive, the reference answers every per-cell overlap through lib/tcm-memo (a C++ hash map memoizing the Sankoff
 * median/cost of a pair of ambiguity sets, lib/tcm-memo/ffi/memoized-tcm/costMatrix_2d.cpp:131-267).  Here the
 * overlap is a device function and the memo a per-pair dense table in shared memory / HBM (do_explicit.cu).
 * sigma: a*a row-major, sigma[s*a + j] = cost between MEDIAN symbol s and INPUT symbol j, gap = symbol a-1; a <= 32. */
typedef struct pcgdo_memo pcgdo_memo;
int  pcgdo_memo_create(pcgdo_ctx *ctx, const unsigned int *sigma, size_t a, pcgdo_memo **out);
void pcgdo_memo_destroy(pcgdo_ctx *ctx, pcgdo_memo *memo);
/* The metric representation PCG derives from a (factored) symbol-change matrix — `dynamicMetadata`,
 * Bio/Metadata/Dynamic/Internal.hs:286-300 over `diagnoseTcm`, Data/TCM/Internal.hs:506-524 — decides the overlap function
 * of the a > 8 path (`retreivePairwiseTCM`, Data/MetricRepresentation.hs:84-91):
 *   PCGDO_METRIC_EXPLICIT  memoized Sankoff overlap of the matrix (tcm-memo)        sigma required
 *   PCGDO_METRIC_DISCRETE  discreteMetricPairwiseLogic  (:116-128)                  sigma may be NULL
 *   PCGDO_METRIC_L1        firstLinearNormPairwiseLogic (:166-190), as coded: a set is read as the interval lowest..highest
 *                          set symbol and the median of [lb, ub], lb < ub, is the symbols lb..ub-1   sigma may be NULL
 *   PCGDO_METRIC_AUTO      diagnose sigma like PCG: |i-j| -> L1, 0/1 -> discrete, else explicit (factor the matrix first)
 * Alphabets of up to PCGDO_MAX_ALPHABET symbols: beyond 32 an element is W = ceil(a / 32) consecutive 32-bit words (bit i
 * of the set = bit i % 32 of word i / 32, gap = symbol a-1; elements order as unsigned naturals, most significant word
 * first), in pcgdo_batch32 and in forests alike; offsets, lengths and capacities keep counting ELEMENTS. */
#define PCGDO_MAX_ALPHABET     512
#define PCGDO_METRIC_AUTO     (-1)
#define PCGDO_METRIC_EXPLICIT   0
#define PCGDO_METRIC_DISCRETE   1
#define PCGDO_METRIC_L1         2
int  pcgdo_memo_create_metric(pcgdo_ctx *ctx, const unsigned int *sigma, size_t a, int structure, pcgdo_memo **out);
int  pcgdo_memo_structure(const pcgdo_memo *memo);
uint32_t pcgdo_memo_alphabet(const pcgdo_memo *memo);
/* n independent overlap lookups: median[k], cost[k] = Sankoff median / cost of the sets e1[k], e2[k]. */
int  pcgdo_overlap2d_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, size_t n, const uint32_t *e1, const uint32_t *e2,
                                  uint32_t *median, uint32_t *cost, void *cuda_stream);
int  pcgdo_overlap2d_batch_host(pcgdo_ctx *ctx, const pcgdo_memo *memo, size_t n, const uint32_t *e1, const uint32_t *e2,
                                uint32_t *median, uint32_t *cost);
/* The three Haskell dialects of pcgdo_metric_batch_* over the explicit matrix (same batch layout and outputs). */
int  pcgdo_explicit_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, int dialect, const pcgdo_batch32 *b, void *cuda_stream);
int  pcgdo_explicit_batch_host(pcgdo_ctx *ctx, const pcgdo_memo *memo, int dialect, const pcgdo_batch32 *b);

/* n independent lookups over alphabets of up to PCGDO_MAX_ALPHABET symbols and over two OR THREE sets (nsets = 2 / 3;
 * e3 ignored for 2): set k of array ex is the W = ceil(a / 32) words ex[k*W .. k*W+W), median[k*W ..] likewise.  The
 * three-set form is costMatrix_3d.cpp:168-211 (computeCostMedian over three keys).  One warp per lookup. */
int  pcgdo_overlap_sets_batch_device(pcgdo_ctx *ctx, const pcgdo_memo *memo, int nsets, size_t n, const uint32_t *e1,
                                     const uint32_t *e2, const uint32_t *e3, uint32_t *median, uint32_t *cost, void *cuda_stream);
int  pcgdo_overlap_sets_batch_host(pcgdo_ctx *ctx, const pcgdo_memo *memo, int nsets, size_t n, const uint32_t *e1,
                                   const uint32_t *e2, const uint32_t *e3, uint32_t *median, uint32_t *cost);

/* Reference-compatible tcm-memo symbols (lib/tcm-memo/ffi/memoized-tcm/costMatrixWrapper.h:12-31,
 * dynamicCharacterOperations.h; bound at lib/tcm-memo/src/Data/TCM/Memoized/FFI.hsc:49-69): one lookup per call
 * through the GPU, alphabets up to PCGDO_MAX_ALPHABET (elements are ceil(a / 64) uint64 words, as the reference packs
 * them).  getCostAndMedian2D / 3D free and re-allocate retElem->element like the reference
 * (costMatrix_2d.cpp:156-166, costMatrix_3d.cpp:142-145); retElem may be NULL for the 3D lookup (cost only). */
typedef struct dcElement_t { size_t alphSize; uint64_t *element; } dcElement_t;
typedef void *costMatrix_p;
costMatrix_p matrixInit(size_t alphSize, unsigned int *tcm);
void         matrixDestroy(costMatrix_p untyped_ptr);
unsigned int getCostAndMedian2D(dcElement_t *elem1, dcElement_t *elem2, dcElement_t *retElem, costMatrix_p tcm);
unsigned int getCostAndMedian3D(dcElement_t *elem1, dcElement_t *elem2, dcElement_t *elem3, dcElement_t *retElem,
                                costMatrix_p tcm);

/* Finer tunables: widest diagonals-per-lane variant the warp kernels must accommodate (2,4,8,16,32; pairs
 * needing more go to the general-geometry kernel) and the per-warp scratch arena in 32-bit words (0 = derive
 * from max_len). */
int pcgdo_configure_ex(pcgdo_ctx *ctx, int max_b, size_t arena_words);

/* ---------------------------------------------------------------- post-order tree scoring
 * Replaces the post-order sweep that calls the pairwise DO once per internal node and dynamic character
 * (Bio/Graph/PhylogeneticDAG/Postorder.hs:98-158, DO/Internal.hs:109-140): all nodes of equal height
 * across all candidate trees and characters are aligned in one launch; nodes keep their ungapped median
 * strings resident in HBM.
 *   n_taxa leaves have node ids 0 .. n_taxa-1 (shared by all trees), internal node x of a tree has id
 *   n_taxa + x; children[(t*(n_taxa-1) + x)*2 + {0,1}] are the two child ids of internal node x of tree t.
 *   node_cap = byte capacity of every node's median string. */
typedef struct pcgdo_forest pcgdo_forest;
int  pcgdo_forest_create(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars, uint32_t node_cap,
                         const int32_t *children, pcgdo_forest **out);
void pcgdo_forest_destroy(pcgdo_ctx *ctx, pcgdo_forest *forest);
/* HOST buffers: character c of taxon v = elems[off[v*n_chars+c] .. +len[v*n_chars+c]) (no gap elements).
 * len = PCGDO_MISSING marks a Missing character (a taxon without data for that character): aligning it with anything
 * costs nothing and yields the other operand (`handleMissingCharacter`, DO/Pairwise/Internal.hs:247-259); a node whose
 * two children are Missing — or both empty once ungapped — is Missing itself (get_node reports that length), and an
 * alignment with ONE empty operand is free, as in the reference (`directOptimization`, :224-233). */
#define PCGDO_MISSING 0xffffffffu
int  pcgdo_forest_set_leaves(pcgdo_ctx *ctx, pcgdo_forest *forest, const uint8_t *elems, const uint64_t *off,
                             const uint32_t *len);
/* One sweep: every internal node's local cost, total cost (local + children, DO/Internal.hs:133-137) and
 * ungapped median; per-tree cost = sum over this forest's characters of the root total. */
int  pcgdo_forest_score(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, pcgdo_forest *forest, void *cuda_stream);
void *pcgdo_forest_tree_cost_ptr(pcgdo_forest *forest);              /* DEVICE int64[n_trees]          */
int  pcgdo_forest_tree_costs(pcgdo_ctx *ctx, pcgdo_forest *forest, int64_t *host_out);
int  pcgdo_forest_copy_tree_costs(pcgdo_ctx *ctx, pcgdo_forest *forest, void *dst_device, void *cuda_stream);
int  pcgdo_forest_get_node(pcgdo_ctx *ctx, pcgdo_forest *forest, uint32_t tree, uint32_t character, uint32_t node,
                           uint8_t *median_out, uint32_t *len_out, int32_t *local_cost, int64_t *total_cost);
/* Large alphabets (a > 8): the same scheduler over 32-bit elements, every alignment by the Haskell dialects of
 * pcgdo_metric_batch_* / pcgdo_explicit_batch_* (memo = NULL: discrete metric over `alphabet` symbols; dialect 2 =
 * unboxedUkkonenFullSpaceDO, what selectDynamicMetric picks for a > 8).  score / score_rerooted / try_leaf /
 * get_node take such a forest with tcm = NULL; strings are then uint32 elements (get_node, try_leaf: cast). */
int  pcgdo_forest_create32(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars, uint32_t node_cap,
                           const int32_t *children, const pcgdo_memo *memo, int alphabet, int dialect, pcgdo_forest **out);
int  pcgdo_forest_set_leaves32(pcgdo_ctx *ctx, pcgdo_forest *forest, const uint32_t *elems, const uint64_t *off,
                               const uint32_t *len);
uint32_t pcgdo_forest_levels(const pcgdo_forest *forest);
/* Re-rooting of dynamic characters (Bio/Graph/PhylogeneticDAG/DynamicCharacterRerooting.hs:70-420, trees only): after
 * the post-order, directed subtree data for every internal node (a pre-order sweep by depth) and one cost-only
 * alignment per unrooted edge; per-tree cost = sum over characters of the minimum over the 2*n_taxa-3 edges (each
 * dynamic character picks its own root edge).  enable_rerooting grows the node storage to 3 slots per internal
 * node.  get_edges: (u, v) node ids of one tree's edges (edge 0 = the original root edge) and, for one character,
 * the total / local cost of rooting on each.  keep_edge_data = 1 additionally keeps every edge's datum (the median of
 * rooting the tree there) for Wagner-build edge trials (app/pcg/PCG/Command/Build/Evaluate.hs:494-527, tryEdge):
 * pcgdo_forest_try_leaf aligns a new leaf (HOST strings, one per character) with every edge datum of every tree, cost
 * only, and returns delta[tree * n_edges + edge] = sum over characters of the local cost (HOST int64) — the cost
 * increase of attaching the leaf to that edge. */
int  pcgdo_forest_enable_rerooting(pcgdo_ctx *ctx, pcgdo_forest *forest, int keep_edge_data);
int  pcgdo_forest_score_rerooted(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, pcgdo_forest *forest, void *cuda_stream);
int  pcgdo_forest_get_edges(pcgdo_ctx *ctx, pcgdo_forest *forest, uint32_t tree, uint32_t character, int32_t *edge_uv,
                            int64_t *edge_total, int32_t *edge_local);
uint32_t pcgdo_forest_edges(const pcgdo_forest *forest);
int  pcgdo_forest_try_leaf(pcgdo_ctx *ctx, const pcgdo_tcm *tcm, pcgdo_forest *forest, const uint8_t *elems,
                           const uint64_t *off, const uint32_t *len, int64_t *delta_out, void *cuda_stream);
uint64_t pcgdo_forest_alignments(const pcgdo_forest *forest);
/* Candidate-tree turnover: replace all n_trees topologies (same layout as `children` at creation) over the resident
 * leaves.  The new trees are validated and levelled ON THE DEVICE into the buffers allocated at creation (sized for any
 * topology over these taxa): nothing is allocated, the next score / score_rerooted sweeps the new candidates.  This is
 * what a search does between sweeps (app/pcg/PCG/Command/Build/Evaluate.hs:494-527 yields new topologies over fixed leaves). */
int  pcgdo_forest_set_trees(pcgdo_ctx *ctx, pcgdo_forest *forest, const int32_t *children);
/* DP cells of the reference's band over the alignments of the last post-order sweep (8-bit forests): the numerator of the
 * sweep's GCUPS / roofline figures. */
int  pcgdo_forest_cells(pcgdo_ctx *ctx, pcgdo_forest *forest, uint64_t *cells_out);

/* ---------------------------------------------------------------- pre-order pass: final ("implied") alignments
 * Replaces directOptimizationPreorder / initializeRoot / updateFromParent / deriveImpliedAlignment / lexicallyDisambiguate
 * (lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization/Internal.hs:148-286) for rooted binary trees over
 * dense-TCM alphabets (one byte per ambiguity set): from the post-order CONTEXTS of every node — (median, left, right)
 * strings with gaps, a leaf's being (s, s, s) — to the final alignment of every node, all of the root's length, plus its
 * single (lexical) disambiguation.  Contexts are uploaded once and stay on the device; the descent is one launch per depth
 * over all trees x characters (do_preorder.cu).  Node ids and `children` as in pcgdo_forest_create.
 *   set_contexts  HOST arrays: node v of column (tree t, character c) = med / left / right [off[i] .. off[i] + len[i]),
 *                 i = (t * n_chars + c) * (2 * n_taxa - 1) + v; len[i] = PCGDO_MISSING for a Missing character
 *   run           alphabet = symbols including the gap (<= 8); PCGDO_ERR_OVERFLOW when an edge reaches the reference's
 *                 "Impossible happened in 'deriveImpliedAlignment'" (the contexts do not belong to the tree)
 *   get           one node's final alignment, pcgdo_preorder_length bytes per plane (any output may be NULL)
 *   device_ptr    the resident result: plane p (0 median, 1 left, 2 right, 3 single) of node v of column col starts at
 *                 base + off[col] + (v * 4 + p) * round4(len[col]), off being the running sum of
 *                 (2 * n_taxa - 1) * 4 * round4(len) over the non-Missing columns before col */
typedef struct pcgdo_preorder pcgdo_preorder;
int  pcgdo_preorder_create(pcgdo_ctx *ctx, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars, const int32_t *children,
                           pcgdo_preorder **out);
void pcgdo_preorder_destroy(pcgdo_ctx *ctx, pcgdo_preorder *pre);
int  pcgdo_preorder_set_contexts(pcgdo_ctx *ctx, pcgdo_preorder *pre, const uint8_t *med, const uint8_t *left,
                                 const uint8_t *right, size_t total_bytes, const uint64_t *off, const uint32_t *len);
int  pcgdo_preorder_run(pcgdo_ctx *ctx, pcgdo_preorder *pre, int alphabet, void *cuda_stream);
uint32_t pcgdo_preorder_length(const pcgdo_preorder *pre, uint32_t tree, uint32_t character);
int  pcgdo_preorder_get(pcgdo_ctx *ctx, pcgdo_preorder *pre, uint32_t tree, uint32_t character, uint32_t node,
                        uint8_t *med_out, uint8_t *left_out, uint8_t *right_out, uint8_t *single_out, int *missing_out);
void *pcgdo_preorder_device_ptr(pcgdo_preorder *pre);

/* ---------------------------------------------------------------- several GPUs in one process
 * north_star item (4) as a LIBRARY feature: a caller that is one process (PCG is) hands the library a device list; forests
 * are sharded by CHARACTER BLOCKS across the devices (every device keeps all node strings of its characters resident and
 * runs all levels locally — no traffic during the sweep), and the per-tree sums int64[n_trees] are added across devices
 * after the last level.  When there are fewer characters than devices the trees are split instead (no reduction at all).
 * reduce_mode: PCGDO_REDUCE_HOST  one 8 KB device-to-host copy per device, added on the host;
 *              PCGDO_REDUCE_PEER  one kernel on the first device reads the other devices' sums over NVLink (peer access);
 *              PCGDO_REDUCE_NCCL  ncclAllReduce(sum, int64) over a single-process communicator (libnccl.so.2 is loaded at
 *                                 run time; PCGDO_ERR_UNSUPPORTED when it cannot be). */
typedef struct pcgdo_multi pcgdo_multi;
typedef struct pcgdo_mforest pcgdo_mforest;
typedef struct pcgdo_mtcm pcgdo_mtcm;
enum { PCGDO_REDUCE_HOST = 0, PCGDO_REDUCE_PEER = 1, PCGDO_REDUCE_NCCL = 2 };
int  pcgdo_multi_create(pcgdo_multi **out, const int *devices, int n_devices);
void pcgdo_multi_destroy(pcgdo_multi *m);
int  pcgdo_multi_devices(const pcgdo_multi *m);
pcgdo_ctx  *pcgdo_multi_ctx(pcgdo_multi *m, int i);
const char *pcgdo_multi_last_error(const pcgdo_multi *m);
int  pcgdo_multi_configure(pcgdo_multi *m, int pairs_per_warp, int warps_per_cta, int max_len, int max_b);
int  pcgdo_multi_tcm_create(pcgdo_multi *m, const unsigned int *sigma, size_t a, pcgdo_mtcm **out);
void pcgdo_multi_tcm_destroy(pcgdo_multi *m, pcgdo_mtcm *t);
int  pcgdo_mforest_create(pcgdo_multi *m, uint32_t n_trees, uint32_t n_taxa, uint32_t n_chars, uint32_t node_cap,
                          const int32_t *children, pcgdo_mforest **out);
void pcgdo_mforest_destroy(pcgdo_multi *m, pcgdo_mforest *f);
/* leaves as pcgdo_forest_set_leaves (all n_chars characters; each device takes its block) */
int  pcgdo_mforest_set_leaves(pcgdo_multi *m, pcgdo_mforest *f, const uint8_t *elems, const uint64_t *off, const uint32_t *len);
int  pcgdo_mforest_set_trees(pcgdo_multi *m, pcgdo_mforest *f, const int32_t *children);
/* one sweep on all devices at once + the reduction; tree_costs: HOST int64[n_trees]; rerooted != 0 adds the re-rooting pass
 * (pcgdo_mforest_enable_rerooting first).  wall_ms (may be NULL): host wall time of the call; sweep_ms (may be NULL): the
 * slowest device's sweep by its own CUDA events; reduce_ms (may be NULL): host wall time of the reduction step alone. */
int  pcgdo_mforest_enable_rerooting(pcgdo_multi *m, pcgdo_mforest *f);
int  pcgdo_mforest_score(pcgdo_multi *m, const pcgdo_mtcm *tcm, pcgdo_mforest *f, int rerooted, int reduce_mode,
                         int64_t *tree_costs, double *wall_ms, float *sweep_ms, double *reduce_ms);
/* the character block [lo, hi) (or tree block when the forest is split by trees: *by_trees = 1) of device i */
int  pcgdo_mforest_block(const pcgdo_mforest *f, int i, uint32_t *lo, uint32_t *hi, int *by_trees);
/* Pair batches over several devices: the pairs are split evenly, each device runs the pipelined host entry point on its
 * share from its own host thread (weak sharding, no collective).  Same pcgdo_batch contract as pcgdo_align2d_batch_host. */
int  pcgdo_multi_align2d_batch_host(pcgdo_multi *m, const pcgdo_mtcm *tcm, const pcgdo_batch *b);

/* Measured integer instruction-issue rate of this GPU in lane-instructions/s (roofline denominator of the
 * DP fill; mode 0: ALU-pipe mix, 1: VIADDMNMX + IMAD, one per pipe). */
int pcgdo_int32_peak(pcgdo_ctx *ctx, int mode, double *ops_per_s, float *ms_out);

/* Timing of the kernels of the most recent *_device / *_host call on this context, measured with
 * CUDA events on the launching stream (ms); also the number of kernel launches it made. */
int pcgdo_last_kernel_ms(pcgdo_ctx *ctx, float *ms, int *launches);

#ifdef __cplusplus
}
#endif
#endif /* PCG_DO_B200_H */
