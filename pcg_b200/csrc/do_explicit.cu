// do_explicit.cu — the Haskell DO dialects over an EXPLICIT symbol-change matrix (alphabets of 9..32 symbols),
// sm_100a.  Replaces, for this path, the memoized Sankoff lookups of lib/tcm-memo:
//   getMedianAndCost2D      lib/tcm-memo/src/Data/TCM/Memoized/FFI.hsc:97-112        (one FFI call per DP cell)
//   CostMatrix_2d::getSetCostMedian / computeCostMedian / findDistance
//                           lib/tcm-memo/ffi/memoized-tcm/costMatrix_2d.cpp:131-267   (hash map keyed on the pair)
//   overlap / overlap2      lib/core/data-structures/src/Bio/Metadata/Overlap.hs:56-86, 136-142
// and runs the same three dialects as do_metric.cu (S1 full, S1 Ukkonen, S2 Ukkonen; file references there).
//
// Memoization becomes a PER-PAIR DENSE TABLE: the distinct ambiguity sets of the two strings are numbered
// (warp-parallel open-addressing hash), the Sankoff cost/median of every (distinct longer, distinct lesser)
// combination is computed once — from per-element profiles m_x[s] = min_{j in x} sigma[s][j] — and stored as
// the same folded int16 value the dense-table kernel uses, 4*(cost(x,y) - cost(x,gap) - cost(gap,y)) - 1 (+1 when
// S2 and the median is exactly the gap).  The table sits in shared memory when it is small (single-symbol
// strings over 21 amino acids: 22 x 22 entries) and in the warp's global scratch otherwise.  The fill is then
// the dense-table recurrence: one address IMAD, one load, two VIADDMNMX, one LOP3, two IMAD per cell.
// PARITY UNPINNED for the dialect logic (see do_metric.cu); the overlap itself is pinned against the compiled
// reference tcm-memo (tests/golden/tcm_memo_golden.json).
#include "do_metric.cuh"
#include "do_explicit.h"
#include "do_pack16.cuh"

namespace pcgdo {

// Sankoff median / cost of two ambiguity sets straight from sigma (shared memory), costMatrix_2d.cpp:175-267.
__device__ __forceinline__ void ov_explicit(const uint32_t *sg, int a, uint32_t x, uint32_t y, uint32_t &med, uint32_t &cost)
{
    uint32_t best = 0xffffffffu, bits = 0;
    for (int s = 0; s < a; ++s) {
        const uint32_t *row = sg + s * a;
        uint32_t dx = 0xffffffffu, dy = 0xffffffffu;
        for (uint32_t t = x; t; t &= t - 1) dx = min(dx, row[__ffs(t) - 1]);
        for (uint32_t t = y; t; t &= t - 1) dy = min(dy, row[__ffs(t) - 1]);
        const uint32_t c = dx + dy;                      // unsigned wrap for empty sets, as the reference
        if (c < best) { best = c; bits = 1u << s; }
        else if (c == best) bits |= 1u << s;
    }
    med = bits; cost = best;
}

// firstLinearNormPairwiseLogic (lib/core/tcm/src/Data/MetricRepresentation.hs:166-190) over Ranged AmbiguityGroup
// (Bio/Character/Encodable/Dynamic/AmbiguityGroup.hs:154-172, lib/utility/src/Data/Range.hs:100-190): each set is read as
// the interval lowest..highest set symbol; overlapping intervals intersect at cost 0, disjoint ones give the closed gap
// between them at its length; fromRange of lb < ub sets symbols lb .. ub-1 ((2^ub - 1) xor (2^lb - 1)), of lb == ub the
// one symbol.  Returns the median as an inclusive symbol range [mlo, mhi] (empty when mlo > mhi).
// PINNED: the reference's L1 script costs 3483 / 21851 (tests/golden/custom_alphabet_scripts.json).
__device__ __forceinline__ uint32_t l1_ranges(int a, int lo0, int hi0, int lo1, int hi1, int &mlo, int &mhi)
{
    const bool isect = lo1 <= hi0 && hi1 >= lo0;
    int nl, nu;
    if (isect) { nl = max(lo0, lo1); nu = min(hi0, hi1); }
    else       { nl = min(hi0, hi1); nu = max(lo0, lo1); }
    if (nu == nl) { mlo = nl; mhi = nl < a ? nl : -1; }
    else          { mlo = nl; mhi = min(nu - 1, a - 1); }
    return isect ? 0u : (uint32_t)(nu - nl);
}
__device__ __forceinline__ void l1_lohi(uint32_t x, int a, int &lo, int &hi)
{
    const int l = x ? __ffs(x) - 1 : a, h = x ? 31 - __clz(x) : 0;
    lo = min(l, h); hi = max(l, h);
}
__device__ __forceinline__ void ov_l1(int a, uint32_t x, uint32_t y, uint32_t &med, uint32_t &cost)
{
    int lo0, hi0, lo1, hi1, mlo, mhi;
    l1_lohi(x, a, lo0, hi0);
    l1_lohi(y, a, lo1, hi1);
    cost = l1_ranges(a, lo0, hi0, lo1, hi1, mlo, mhi);
    med = (mhi < mlo) ? 0u : (uint32_t)((((1ull << (mhi + 1)) - 1ull) ^ ((1ull << mlo) - 1ull)) & 0xffffffffull);
}

// ---- alphabets of more than 32 symbols: an element is W = ceil(a / 32) consecutive words, bit i of the set = bit i % 32
// of word i / 32; offsets and lengths keep counting elements.  Ordering of elements (measureCharacters' tie-break) is
// that of unsigned naturals, most significant word first, like bv-little's Ord.  Plain (generic) loads: an operand may
// be the gap element staged in shared memory.
__device__ __forceinline__ int w_cmp(const uint32_t *x, const uint32_t *y, int W)
{
    for (int w = W - 1; w >= 0; --w) {
        const uint32_t a = x[w], b = y[w];
        if (a != b) return a < b ? -1 : 1;
    }
    return 0;
}
__device__ __forceinline__ bool w_eq(const uint32_t *x, const uint32_t *y, int W)
{
    for (int w = 0; w < W; ++w) if (x[w] != y[w]) return false;
    return true;
}
__device__ __forceinline__ bool w_zero(const uint32_t *x, int W)
{
    for (int w = 0; w < W; ++w) if (x[w]) return false;
    return true;
}
__device__ __forceinline__ bool w_test(const uint32_t *x, int bit) { return (x[bit >> 5] >> (bit & 31)) & 1u; }
__device__ __forceinline__ uint32_t w_hash(const uint32_t *x, int W)
{
    uint32_t h = 0x811C9DC5u;
    for (int w = 0; w < W; ++w) h = (h ^ x[w]) * 0x01000193u;
    return (h * 0x9E3779B1u) >> 7;
}
__device__ __forceinline__ void w_lohi(const uint32_t *x, int W, int a, int &lo, int &hi)
{
    int l = a, h = 0;
    bool any = false;
    for (int w = 0; w < W; ++w) {
        const uint32_t v = x[w];
        if (v) {
            if (!any) { l = 32 * w + __ffs(v) - 1; any = true; }
            h = 32 * w + 31 - __clz(v);
        }
    }
    lo = min(l, h); hi = max(l, h);
}
// min over the set symbols j of x of row[j]
__device__ __forceinline__ uint32_t w_rowmin(const uint32_t *row, const uint32_t *x, int W)
{
    uint32_t mn = 0xffffffffu;
    for (int w = 0; w < W; ++w)
        for (uint32_t t = x[w]; t; t &= t - 1) mn = min(mn, row[32 * w + __ffs(t) - 1]);
    return mn;
}
// the overlap of two W-word sets; med (W words, may be NULL) receives the median.  kind: 0 discrete, 1 L1, 2 explicit.
__device__ __forceinline__ uint32_t w_overlap(int kind, const uint32_t *sg, int a, int W, const uint32_t *x, const uint32_t *y,
                                              uint32_t *med)
{
    if (kind == 0) {
        bool inter = false;
        for (int w = 0; w < W; ++w) inter = inter || (x[w] & y[w]);
        if (med) for (int w = 0; w < W; ++w) med[w] = inter ? (x[w] & y[w]) : (x[w] | y[w]);
        return inter ? 0u : 1u;
    }
    if (kind == 1) {
        int lo0, hi0, lo1, hi1, mlo, mhi;
        w_lohi(x, W, a, lo0, hi0);
        w_lohi(y, W, a, lo1, hi1);
        const uint32_t c = l1_ranges(a, lo0, hi0, lo1, hi1, mlo, mhi);
        if (med)
            for (int w = 0; w < W; ++w) {
                const int b0 = max(mlo, 32 * w) - 32 * w, b1 = min(mhi, 32 * w + 31) - 32 * w;
                med[w] = (b1 < b0) ? 0u : (uint32_t)((((1ull << (b1 + 1)) - 1ull) ^ ((1ull << b0) - 1ull)) & 0xffffffffull);
            }
        return c;
    }
    uint32_t best = 0xffffffffu;
    if (med) for (int w = 0; w < W; ++w) med[w] = 0u;
    for (int s = 0; s < a; ++s) {
        const uint32_t *row = sg + (size_t)s * a;
        const uint32_t c = w_rowmin(row, x, W) + w_rowmin(row, y, W);
        if (c < best) {
            best = c;
            if (med) { for (int w = 0; w < W; ++w) med[w] = 0u; med[s >> 5] = 1u << (s & 31); }
        } else if (c == best && med) med[s >> 5] |= 1u << (s & 31);
    }
    return best;
}

__global__ void overlap2d_kernel(const uint32_t *sigma, int a, size_t n, const uint32_t *e1, const uint32_t *e2,
                                 uint32_t *med, uint32_t *cost)
{
    __shared__ uint32_t sg[32 * 32];
    for (int i = threadIdx.x; i < a * a; i += blockDim.x) sg[i] = sigma[i];
    __syncthreads();
    const uint32_t mask = a >= 32 ? 0xffffffffu : ((1u << a) - 1u);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        uint32_t m, c;
        ov_explicit(sg, a, e1[k] & mask, e2[k] & mask, m, c);
        if (med) med[k] = m;
        if (cost) cost[k] = c;
    }
}

cudaError_t launch_overlap2d(const uint32_t *sigma, int a, size_t n, const uint32_t *e1, const uint32_t *e2,
                             uint32_t *med, uint32_t *cost, int sm_count, cudaStream_t st)
{
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)sm_count * 8) blocks = (size_t)sm_count * 8;
    if (blocks == 0) blocks = 1;
    overlap2d_kernel<<<(unsigned)blocks, 256, 0, st>>>(sigma, a, n, e1, e2, med, cost);
    return cudaGetLastError();
}

// One WARP per lookup, any alphabet up to PCGDO_MAX_ALPHABET, two or three sets of W words each: the Sankoff median /
// cost of costMatrix_2d.cpp:175-267 (computeCostMedian / findDistance) and costMatrix_3d.cpp:168-211.  Lanes take the
// candidate median symbols s = lane, lane + 32, ...; the median word j is the ballot of "symbol 32 j + lane attains the
// minimum".  Empty sets wrap like the reference's unsigned arithmetic.
__global__ void overlap_sets_kernel(const uint32_t *sigma, int a, int W, int nsets, size_t n, const uint32_t *e1,
                                    const uint32_t *e2, const uint32_t *e3, uint32_t *med, uint32_t *cost)
{
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t k = warp; k < n; k += nwarps) {
        const uint32_t *x = e1 + k * W, *y = e2 + k * W, *z = nsets > 2 ? e3 + k * W : nullptr;
        auto cost_of = [&](int s) {
            const uint32_t *row = sigma + (size_t)s * a;
            uint32_t c = w_rowmin(row, x, W) + w_rowmin(row, y, W);
            if (z) c += w_rowmin(row, z, W);
            return c;
        };
        uint32_t best = 0xffffffffu;
        for (int s = lane; s < a; s += 32) best = min(best, cost_of(s));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
        for (int j = 0; j < W; ++j) {
            const int s = 32 * j + lane;
            const unsigned bits = __ballot_sync(0xffffffffu, s < a && cost_of(s) == best);
            if (med && lane == 0) med[k * W + j] = bits;
        }
        if (cost && lane == 0) cost[k] = best;
    }
}

cudaError_t launch_overlap_sets(const uint32_t *sigma, int a, int nsets, size_t n, const uint32_t *e1, const uint32_t *e2,
                                const uint32_t *e3, uint32_t *med, uint32_t *cost, int sm_count, cudaStream_t st)
{
    size_t blocks = (n + 7) / 8;
    if (blocks > (size_t)sm_count * 8) blocks = (size_t)sm_count * 8;
    if (blocks == 0) blocks = 1;
    overlap_sets_kernel<<<(unsigned)blocks, 256, 0, st>>>(sigma, a, (a + 31) / 32, nsets, n, e1, e2, e3, med, cost);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ fill
__device__ __forceinline__ uint32_t x_lds_s16(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t x_ldg_s16(const char *base, uint32_t off)
{
    uint32_t v;
    asm("ld.global.s16 %0, [%1];" : "=r"(v) : "l"(base + off));
    return v;
}

template <int B, bool GLOBAL, bool TAIL>
__device__ __forceinline__ void x_chunk(uint32_t (&w)[B], const uint32_t (&penc)[B], uint32_t (&dw)[B],
                                        uint32_t (&lw)[B / 2], uint32_t (&sw)[B / 2], uint32_t k4, uint32_t km1,
                                        uint32_t k1, const char *gtab, const uint32_t *&lptr, const uint32_t *&sptr,
                                        int n_iter, uint32_t lane_off, int lane)
{
    constexpr int H = B / 2;
    // bytes between the table copies of neighbouring lanes: lane 1's offset (lane 0's is 0 in either mode — testing
    // `lane_off` itself left lane 0 on lane 1's copy and bank: every fetch took two wavefronts)
    const uint32_t lane_step = __shfl_sync(kFullM, lane_off, 1);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (TAIL && r >= n_iter) break;
        const int tt = r % H;
        const uint32_t wl = __shfl_up_sync(kFullM, w[B - 1], 1);
#pragma unroll
        for (int u = 0; u < H; ++u) {
            const int m = 2 * u;
            const uint32_t left = (u == 0) ? wl : w[m > 0 ? m - 1 : 0];
            const uint32_t ad = m_imad(lw[(tt - u + H) % H], k1, sw[(tt + u) % H]);
            const uint32_t sub = GLOBAL ? x_ldg_s16(gtab, ad) : x_lds_s16(ad);
            const uint32_t t1 = __viaddmin_u32(w[m], sub, w[m + 1]);
            const uint32_t mn = __viaddmin_u32(left, 1u, t1);
            const uint32_t wn = (mn & ~3u) | penc[m];
            dw[m] = m_imad(wn, km1, m_imad(dw[m], k4, mn));
            w[m] = wn;
        }
        // Sliding windows by shuffle: the column this lane needs next is the one its right neighbour is about to drop
        // (windows of adjacent lanes are H columns apart), the row it needs next the one its left neighbour drops; only
        // lane 31 / lane 0 read the staged strings.  The per-lane loads this replaces were H-way bank conflicts (lanes H
        // words apart): as many shared-memory wavefronts per iteration as all the table lookups together.
        {
            uint32_t nv = __shfl_down_sync(kFullM, sw[tt], 1) - lane_step;
            if (lane == 31) nv = sptr[r] + lane_off;
            sw[tt] = nv;
        }
        const uint32_t wr = __shfl_down_sync(kFullM, w[0], 1);
#pragma unroll
        for (int u = 0; u < H; ++u) {
            const int m = 2 * u + 1;
            const uint32_t up = (u == H - 1) ? wr : w[m + 1 < B ? m + 1 : B - 1];
            const uint32_t ad = m_imad(lw[(tt - u + H) % H], k1, sw[(tt + u + 1) % H]);
            const uint32_t sub = GLOBAL ? x_ldg_s16(gtab, ad) : x_lds_s16(ad);
            const uint32_t t1 = __viaddmin_u32(w[m], sub, up);
            const uint32_t mn = __viaddmin_u32(w[m - 1], 1u, t1);
            const uint32_t wn = (mn & ~3u) | penc[m];
            dw[m] = m_imad(wn, km1, m_imad(dw[m], k4, mn));
            w[m] = wn;
        }
        {
            uint32_t nl = __shfl_up_sync(kFullM, lw[(tt + 1) % H], 1);
            if (lane == 0) nl = lptr[r];
            lw[(tt + 1) % H] = nl;
        }
    }
    lptr += 16;
    sptr += 16;
}

// Lb / Sb: staged row-offset / column-offset words, entry 0 = the leading gap.
template <int B, bool GLOBAL>
__device__ __noinline__ uint32_t x_fill(const MGeom g, uint32_t k4, uint32_t km1, uint32_t k1, const char *gtab,
                                        const uint32_t *Lb, const uint32_t *Sb, uint32_t *slot, int lane, uint32_t lane_off)
{
    constexpr int H = B / 2;
    uint32_t w[B], penc[B], dw[B], lw[H], sw[H];
    const int qb = lane * B;
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int q = qb + m;
        penc[m] = 1u | ((q < g.pad || q >= g.wb) ? 0x80000000u : 0u);
        asm volatile("" : "+r"(penc[m]));
        w[m] = (q == g.q0) ? (kBias4 + 1u) : kInfW;       // table entry of (leading gap, leading gap) is -1
        dw[m] = 0;
    }
    const int I0 = g.q0 / 2 - lane * H, J0 = -(g.q0 / 2) + lane * H;
#pragma unroll
    for (int u = 0; u < H; ++u) {
        lw[(H - u) % H] = Lb[I0 - u];
        sw[u] = Sb[J0 + u] + lane_off;
    }
    const uint32_t *lptr = Lb + (I0 + 1), *sptr = Sb + (J0 + H);
    uint32_t *out = slot + qb;
    const int nfull = g.T >> 4, rem = g.T & 15;
    for (int c = 0; c < nfull; ++c) {
        x_chunk<B, GLOBAL, false>(w, penc, dw, lw, sw, k4, km1, k1, gtab, lptr, sptr, 16, lane_off, lane);
        m_store<B>(out, dw, 16);
        out += 32 * B;
    }
    if (rem) {
#pragma unroll
        for (int m = 0; m < B; ++m) dw[m] = 0;
        x_chunk<B, GLOBAL, true>(w, penc, dw, lw, sw, k4, km1, k1, gtab, lptr, sptr, rem, lane_off, lane);
        m_store<B>(out, dw, rem);
    }
    const int q_last = (g.sh - 1) - (g.lg - 1) + g.q0, m_last = q_last % B;
    uint32_t v = 0;
#pragma unroll
    for (int m = 0; m < B; ++m)
        if (m == m_last) v = w[m];
    return __shfl_sync(kFullM, v, q_last / B);
}

template <bool GLOBAL>
__device__ __forceinline__ uint32_t x_dispatch(int B, const MGeom &g, uint32_t k4, uint32_t km1, uint32_t k1,
                                               const char *gtab, const uint32_t *Lb, const uint32_t *Sb, uint32_t *slot,
                                               int lane, uint32_t lane_off)
{
    switch (B) {
        case 2:  return x_fill<2, GLOBAL>(g, k4, km1, k1, gtab, Lb, Sb, slot, lane, lane_off);
        case 4:  return x_fill<4, GLOBAL>(g, k4, km1, k1, gtab, Lb, Sb, slot, lane, lane_off);
        case 8:  return x_fill<8, GLOBAL>(g, k4, km1, k1, gtab, Lb, Sb, slot, lane, lane_off);
        case 16: return x_fill<16, GLOBAL>(g, k4, km1, k1, gtab, Lb, Sb, slot, lane, lane_off);
        default: return x_fill<32, GLOBAL>(g, k4, km1, k1, gtab, Lb, Sb, slot, lane, lane_off);
    }
}

// ------------------------------------------------------------------------------------------------ packed fill
// Same scheme as do_linear.cu's fill_pair16: two diagonals per register, VIADDMNMX.U16x2, windowed re-basing.
__device__ __forceinline__ uint32_t x_lds_u16(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t x_ldg_u16(const char *base, uint32_t off)
{
    uint32_t v;
    asm("ld.global.u16 %0, [%1];" : "=r"(v) : "l"(base + off));
    return v;
}

// Iterations unrolled per chunk of the packed fill.  32 diagonals per lane: FOUR — the 8-iteration body was 22 KB of
// SASS; with the code of the other phases (staging, trace, emit, the narrow fills of the doubling schedule) running on
// the same SM it fell out of the 32 KB instruction cache: a quarter of the fill's stall samples on config 4 were
// "no instruction".  A direction word still covers 8 iterations: `phase` says which half of it a 4-iteration chunk is.
__host__ __device__ constexpr int x_chunk_len(int B) { return B <= 8 ? 16 : (B == 16 ? 8 : 4); }

// The H packed updates of an iteration (H / 2 even-diagonal registers, then H / 2 odd ones) are numbered f = 0 .. H-1.
// The two table fetches of update f + D are issued right before the arithmetic of update f, into a ring of D register
// pairs (they depend on the string windows only): every fetch has D updates of arithmetic between issue and use instead
// of none, which is what the 14 .. 16 resident warps of this kernel could not hide (short_scoreboard was its first stall
// reason), and the stream alternates between the ALU pipe and the FMA / LSU pipes.  Measured first on the four-pairs-per-
// warp kernel of do_linear.cu.
template <int B> __host__ __device__ constexpr int x_dist() { return B / 4 >= 4 ? 4 : B / 4; }

// fetches of update g in the frame of iteration tt (g >= H: update g - H of the next iteration)
template <int B, bool GLOBAL>
__device__ __forceinline__ void x_fetch(int tt, int g, const uint32_t (&lw)[B / 2], const uint32_t (&sw)[B / 2], uint32_t k1,
                                        const char *gtab, uint32_t &lo, uint32_t &hi)
{
    constexpr int H = B / 2, Q = H / 2, U = H / 2;
    const int t2 = tt + g / H, gg = g % H, odd = gg >= U ? 1 : 0, u = gg - U * odd;
    const uint32_t a_lo = m_imad(lw[(t2 - u + 2 * H) % H], k1, sw[(t2 + u + odd) % H]);
    const uint32_t a_hi = m_imad(lw[(t2 - u - Q + 2 * H) % H], k1, sw[(t2 + u + Q + odd) % H]);
    lo = GLOBAL ? x_ldg_u16(gtab, a_lo) : x_lds_u16(a_lo);
    hi = GLOBAL ? x_ldg_u16(gtab, a_hi) : x_lds_u16(a_hi);
}
// (a & b) | c as one LOP3 (ptxas splits the C expression in two for about half of the updates)
__device__ __forceinline__ uint32_t x_and_or(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int B, bool GLOBAL, bool TAIL>
__device__ __forceinline__ void x_chunk16(uint32_t (&R)[B / 2], const uint32_t (&penc)[B / 2], uint32_t (&dw)[B / 2],
                                          uint32_t (&lw)[B / 2], uint32_t (&sw)[B / 2], uint32_t (&p_lo)[x_dist<B>()],
                                          uint32_t (&p_hi)[x_dist<B>()], uint32_t k4, uint32_t k1,
                                          uint32_t k64k, const char *gtab, const uint32_t *&lptr, const uint32_t *&sptr,
                                          int n_iter, uint32_t *&out, uint32_t lane_off, int lane, bool phase)
{
    constexpr int H = B / 2, U = H / 2, CH = x_chunk_len(B), D = x_dist<B>();
    constexpr bool HALF = CH < 8;               // the chunk is half a direction word: `phase` = second half
    const uint32_t lane_step = __shfl_sync(kFull, lane_off, 1);       // (see x_chunk)
#pragma unroll
    for (int r = 0; r < CH; ++r) {
        if (TAIL && r >= n_iter) break;
        const int tt = r % H;
        const bool first = HALF ? (r == 0 && !phase) : (r % 8 == 0);       // first iteration of a direction word
        // sliding windows by shuffle (see x_chunk): the new elements, installed below once no pending fetch wants the old
        uint32_t s_new = __shfl_down_sync(kFull, sw[tt], 1) - lane_step;
        if (lane == 31) s_new = sptr[r] + lane_off;
        uint32_t l_new = __shfl_up_sync(kFull, lw[(tt + 1) % H], 1);
        if (lane == 0) l_new = lptr[r];
        uint32_t edge = 0;
#pragma unroll
        for (int f = 0; f < H; ++f) {
            if (f + D == U) sw[tt] = s_new;                 // before the first odd-diagonal fetch of this iteration
            if (f + D == H) lw[(tt + 1) % H] = l_new;       // before the first fetch of the next one
            const uint32_t sub2 = m_imad(p_hi[f % D], k64k, p_lo[f % D]);
            x_fetch<B, GLOBAL>(tt, f + D, lw, sw, k1, gtab, p_lo[(f + D) % D], p_hi[(f + D) % D]);
            uint32_t mn;
            int m;
            if (f < U) {            // even diagonals: anti-diagonal s = 2t
                m = 2 * f;
                if (f == 0) {
                    const uint32_t prev = __shfl_up_sync(kFull, R[H - 1], 1);
                    edge = __funnelshift_l(prev, R[H - 1], 16);
                }
                const uint32_t left = (m == 0) ? edge : R[m > 0 ? m - 1 : 0];
                const uint32_t t1 = __viaddmin_u16x2(R[m], sub2, R[m + 1 < H ? m + 1 : H - 1]);
                mn = __viaddmin_u16x2(left, 0x00010001u, t1);
            } else {                // odd diagonals: anti-diagonal s = 2t + 1
                m = 2 * (f - U) + 1;
                if (f == U) {
                    const uint32_t next = __shfl_down_sync(kFull, R[0], 1);
                    edge = __funnelshift_r(R[0], next, 16);
                }
                const uint32_t up = (m == H - 1) ? edge : R[m + 1 < H ? m + 1 : H - 1];
                const uint32_t t1 = __viaddmin_u16x2(R[m], sub2, up);
                mn = __viaddmin_u16x2(R[m - 1 >= 0 ? m - 1 : 0], 0x00010001u, t1);
            }
            R[m] = x_and_or(mn, 0xfffcfffcu, penc[m]);
            const uint32_t c2 = mn & 0x00030003u;
            if (HALF && r == 0) dw[m] = m_imad(first ? 0u : dw[m], k4, c2);
            else dw[m] = first ? c2 : m_imad(dw[m], k4, c2);
        }
        if (HALF) {
            const int done = (phase ? 4 : 0) + r + 1;                  // iterations of the current direction word so far
            if (TAIL ? (r == n_iter - 1) : (r == CH - 1 && phase)) {
                store_words16<H>(out, dw, done);
                out += 32 * H;
            }
        } else if (r % 8 == 7) {
            store_words16<H>(out, dw, 8);
            out += 32 * H;
        } else if (TAIL && r == n_iter - 1) {
            store_words16<H>(out, dw, (r % 8) + 1);
            out += 32 * H;
        }
    }
    lptr += CH;
    sptr += CH;
    if constexpr (CH % H != 0) { rot16<H, CH % H>(lw); rot16<H, CH % H>(sw); }
}

// returns the last cell in the 32-bit convention of x_fill, or kOverflow16
template <int B, bool GLOBAL>
__device__ __noinline__ uint32_t x_fill16(const MGeom g, uint32_t k4, uint32_t k1, uint32_t k64k, const char *gtab,
                                          const uint32_t *Lb, const uint32_t *Sb, uint32_t *slot, int lane,
                                          uint32_t spread_limit, uint32_t lane_off)
{
    constexpr int H = B / 2;
    uint32_t R[H], penc[H], dw[H], lw[H], sw[H];
    const int qb = lane * B;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        uint32_t pe = 0, wv = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = qb + m + h * H;
            const uint32_t p1 = 1u | ((q < g.pad || q >= g.wb) ? 0x8000u : 0u);
            const uint32_t w1 = (q == g.q0) ? (kTop16 + 1u) : 0x8001u;
            pe |= p1 << (16 * h);
            wv |= w1 << (16 * h);
        }
        penc[m] = pe;
        asm volatile("" : "+r"(penc[m]));
        R[m] = wv;
        dw[m] = 0;
    }
    const int I0 = g.q0 / 2 - lane * H, J0 = -(g.q0 / 2) + lane * H;
#pragma unroll
    for (int u = 0; u < H; ++u) {
        lw[(H - u) % H] = Lb[I0 - u];
        sw[u] = Sb[J0 + u] + lane_off;
    }
    const uint32_t *lptr = Lb + (I0 + 1), *sptr = Sb + (J0 + H);
    uint32_t *out = slot + lane * H;
    uint32_t p_lo[x_dist<B>()], p_hi[x_dist<B>()];
#pragma unroll
    for (int g0 = 0; g0 < x_dist<B>(); ++g0) x_fetch<B, GLOBAL>(0, g0, lw, sw, k1, gtab, p_lo[g0], p_hi[g0]);
    constexpr int CH = x_chunk_len(B), RB = 32 / CH;
    const int nfull = g.T / CH, rem = g.T % CH;
    int total_shift = 0;
    bool ok = true;
    for (int c = 0; c < nfull; ++c) {
        if (c && c % RB == 0) ok = rebase16<H>(R, total_shift, spread_limit) && ok;
        x_chunk16<B, GLOBAL, false>(R, penc, dw, lw, sw, p_lo, p_hi, k4, k1, k64k, gtab, lptr, sptr, CH, out, lane_off, lane, (c & 1) != 0);
    }
    if (rem) {
        if (nfull && nfull % RB == 0) ok = rebase16<H>(R, total_shift, spread_limit) && ok;
        x_chunk16<B, GLOBAL, true>(R, penc, dw, lw, sw, p_lo, p_hi, k4, k1, k64k, gtab, lptr, sptr, rem, out, lane_off, lane, (nfull & 1) != 0);
    } else if (CH < 8 && (nfull & 1)) {
        // the last direction word holds 4 iterations only: flush it
        store_words16<H>(out, dw, 4);
    }
    ok = rebase16<H>(R, total_shift, spread_limit) && ok;
    if (!ok) return kOverflow16;
    const int q_last = (g.sh - 1) - (g.lg - 1) + g.q0, m_last = q_last % B;
    uint32_t v = 0;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        if (m == m_last) v = R[m] & 0xffffu;
        if (m + H == m_last) v = R[m] >> 16;
    }
    v = __shfl_sync(kFull, v, q_last / B);
    const int x4 = (int)(v & 0x7ffcu) - (int)kTop16 - total_shift;
    return (uint32_t)((int)kBias4 + x4) | 1u;
}

template <bool GLOBAL>
__device__ __forceinline__ uint32_t x_dispatch16(int B, const MGeom &g, uint32_t k4, uint32_t k1, uint32_t k64k,
                                                 const char *gtab, const uint32_t *Lb, const uint32_t *Sb, uint32_t *slot,
                                                 int lane, uint32_t spread, uint32_t lane_off)
{
    switch (B) {
        case 4:  return x_fill16<4, GLOBAL>(g, k4, k1, k64k, gtab, Lb, Sb, slot, lane, spread, lane_off);
        case 8:  return x_fill16<8, GLOBAL>(g, k4, k1, k64k, gtab, Lb, Sb, slot, lane, spread, lane_off);
        case 16: return x_fill16<16, GLOBAL>(g, k4, k1, k64k, gtab, Lb, Sb, slot, lane, spread, lane_off);
        default: return x_fill16<32, GLOBAL>(g, k4, k1, k64k, gtab, Lb, Sb, slot, lane, spread, lane_off);
    }
}

__device__ __forceinline__ uint32_t x_emit(const uint32_t *sg, int a, int kind, const uint32_t *pL, const uint32_t *pS, uint32_t gap,
                                           uint32_t mask, const uint32_t *ops, uint32_t nal, bool swapped, uint32_t *om,
                                           uint32_t *o1, uint32_t *o2, int lane, uint32_t *ou = nullptr, uint32_t ucap = 0)
{
    uint32_t base_i = 0, base_j = 0, base_u = 0;
    for (uint32_t c0 = 0; c0 < nal; c0 += 32) {
        const uint32_t c = c0 + lane;
        uint32_t code = 3u;
        if (c < nal) {
            const uint32_t r = nal - 1u - c;
            code = (__ldcg(ops + (r >> 4)) >> (2 * (r & 15))) & 3u;
        }
        const unsigned bl = __ballot_sync(kFullM, code < 2u), bs = __ballot_sync(kFullM, code == 0u || code == 2u);
        const unsigned lt = (1u << lane) - 1u;
        const uint32_t il = base_i + __popc(bl & lt), js = base_j + __popc(bs & lt);
        bool keep = false;
        uint32_t mdv = 0;
        if (c < nal) {
            const uint32_t y = code < 2u ? (__ldg(pL + il) & mask) : 0u;
            const uint32_t x = (code == 0u || code == 2u) ? (__ldg(pS + js) & mask) : 0u;
            uint32_t md, cc;
            const uint32_t xx = x ? x : gap, yy = y ? y : gap;
            if (kind == 0) md = (xx & yy) ? (xx & yy) : (xx | yy);       // discreteMetricPairwiseLogic
            else if (kind == 1) ov_l1(a, xx, yy, md, cc);
            else ov_explicit(sg, a, xx, yy, md, cc);
            if (om) om[c] = md;
            if (o1) o1[c] = swapped ? y : x;
            if (o2) o2[c] = swapped ? x : y;
            keep = md != gap;
            mdv = md;
        }
        if (ou) {       // deleteGaps on the medians: what the parent's alignment consumes
            const unsigned bk = __ballot_sync(kFullM, keep);
            const uint32_t pos = base_u + __popc(bk & lt);
            if (keep && pos < ucap) ou[pos] = mdv;
            base_u += __popc(bk);
        }
        base_i += __popc(bl);
        base_j += __popc(bs);
    }
    return base_u;
}

constexpr int kMaxW = 16;          // alphabets of up to 512 symbols

// x_emit for W-word elements; gapw = the gap element (W words).  Outputs are W words per position.
__device__ __forceinline__ uint32_t x_emit_w(const uint32_t *sg, int a, int W, int kind, const uint32_t *pL, const uint32_t *pS,
                                             const uint32_t *gapw, const uint32_t *ops, uint32_t nal, bool swapped,
                                             uint32_t *om, uint32_t *o1, uint32_t *o2, int lane, uint32_t *ou, uint32_t ucap)
{
    uint32_t base_i = 0, base_j = 0, base_u = 0;
    for (uint32_t c0 = 0; c0 < nal; c0 += 32) {
        const uint32_t c = c0 + lane;
        uint32_t code = 3u;
        if (c < nal) {
            const uint32_t r = nal - 1u - c;
            code = (__ldcg(ops + (r >> 4)) >> (2 * (r & 15))) & 3u;
        }
        const unsigned bl = __ballot_sync(kFullM, code < 2u), bs = __ballot_sync(kFullM, code == 0u || code == 2u);
        const unsigned lt = (1u << lane) - 1u;
        const uint32_t il = base_i + __popc(bl & lt), js = base_j + __popc(bs & lt);
        bool keep = false;
        uint32_t md[kMaxW];
        if (c < nal) {
            const uint32_t *y = code < 2u ? pL + (size_t)il * W : nullptr;
            const uint32_t *x = (code == 0u || code == 2u) ? pS + (size_t)js * W : nullptr;
            if (y && w_zero(y, W)) y = nullptr;
            if (x && w_zero(x, W)) x = nullptr;
            w_overlap(kind, sg, a, W, x ? x : gapw, y ? y : gapw, md);
            keep = false;
            for (int w = 0; w < W; ++w) {
                const uint32_t xv = x ? x[w] : 0u, yv = y ? y[w] : 0u;
                if (om) om[(size_t)c * W + w] = md[w];
                if (o1) o1[(size_t)c * W + w] = swapped ? yv : xv;
                if (o2) o2[(size_t)c * W + w] = swapped ? xv : yv;
                keep = keep || md[w] != gapw[w];
            }
        }
        if (ou) {       // deleteGaps on the medians: what the parent's alignment consumes
            const unsigned bk = __ballot_sync(kFullM, keep);
            const uint32_t pos = base_u + __popc(bk & lt);
            if (keep && pos < ucap)
                for (int w = 0; w < W; ++w) ou[(size_t)pos * W + w] = md[w];
            base_u += __popc(bk);
        }
        base_i += __popc(bl);
        base_j += __popc(bs);
    }
    return base_u;
}

// x_number for W-word elements: the hash key of a slot is 1 + the position of the first string element that claimed it;
// el[r] = that position.
__device__ __forceinline__ uint32_t x_number_w(const uint32_t *str, int len, int W, uint32_t H, uint32_t *hk, uint32_t *hv,
                                               uint32_t *el, uint32_t *idx, int lane)
{
    for (uint32_t i = lane; i < H; i += 32) hk[i] = 0;
    __syncwarp();
    for (int i = lane; i < len; i += 32) {
        const uint32_t *e = str + (size_t)i * W;
        if (w_zero(e, W)) continue;
        uint32_t h = w_hash(e, W) & (H - 1);
        for (;;) {
            const uint32_t old = atomicCAS(hk + h, 0u, (uint32_t)i + 1u);
            if (old == 0u || w_eq(str + (size_t)(old - 1u) * W, e, W)) break;
            h = (h + 1) & (H - 1);
        }
    }
    __syncwarp();
    uint32_t D = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < H; base += 32) {
        const uint32_t k = __ldcg(hk + base + lane);
        const unsigned bal = __ballot_sync(kFullM, k != 0u);
        if (k) {
            const uint32_t r = D + __popc(bal & lt);
            hv[base + lane] = r + 2u;
            el[r] = k - 1u;
        }
        D += __popc(bal);
    }
    __syncwarp();
    for (int i = lane; i < len; i += 32) {
        const uint32_t *e = str + (size_t)i * W;
        uint32_t v = 0;
        if (!w_zero(e, W)) {
            uint32_t h = w_hash(e, W) & (H - 1);
            while (!w_eq(str + (size_t)(__ldcg(hk + h) - 1u) * W, e, W)) h = (h + 1) & (H - 1);
            v = hv[h];
        }
        idx[i] = v;
    }
    __syncwarp();
    return D;
}

// fewer fill variants = smaller instruction footprint when every warp walks through the doubling schedule
__device__ __forceinline__ int x_choose_b(int wb, uint32_t skip)
{
    int B = m_choose_b(wb);
    while (B && B < 32 && (skip & (uint32_t)B)) B *= 2;
    return B;
}

__host__ __device__ inline uint32_t x_hash_size(uint32_t n)
{
    uint32_t h = 32;
    while (h < 2u * n + 2u) h <<= 1;
    return h;
}
__device__ __forceinline__ uint32_t x_hash(uint32_t e) { return (e * 0x9E3779B1u) >> 7; }

// number the distinct elements of one string: idx[i] = 2 + rank of str[i] among the occupied hash slots
// (0 is the padding entry, 1 the leading gap); elements with no bit set map to the padding entry.
__device__ __forceinline__ uint32_t x_number(const uint32_t *str, int len, uint32_t mask, uint32_t H, uint32_t *hk,
                                             uint32_t *hv, uint32_t *el, uint32_t *idx, int lane)
{
    for (uint32_t i = lane; i < H; i += 32) hk[i] = 0;
    __syncwarp();
    for (int i = lane; i < len; i += 32) {
        const uint32_t e = __ldg(str + i) & mask;
        if (!e) continue;
        uint32_t h = x_hash(e) & (H - 1);
        for (;;) {
            const uint32_t old = atomicCAS(hk + h, 0u, e);
            if (old == 0u || old == e) break;
            h = (h + 1) & (H - 1);
        }
    }
    __syncwarp();
    uint32_t D = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < H; base += 32) {
        const uint32_t k = __ldcg(hk + base + lane);       // written by atomics: bypass L1
        const unsigned bal = __ballot_sync(kFullM, k != 0u);
        if (k) {
            const uint32_t r = D + __popc(bal & lt);
            hv[base + lane] = r + 2u;
            el[r] = k;
        }
        D += __popc(bal);
    }
    __syncwarp();
    for (int i = lane; i < len; i += 32) {
        const uint32_t e = __ldg(str + i) & mask;
        uint32_t v = 0;
        if (e) {
            uint32_t h = x_hash(e) & (H - 1);
            while (__ldcg(hk + h) != e) h = (h + 1) & (H - 1);
            v = hv[h];
        }
        idx[i] = v;
    }
    __syncwarp();
    return D;
}

// Stage one fill's row offsets (bytes into the pair's table) and column offsets (+ the table's shared address), returning
// the indel costs folded out of the cell values.  Its own function: inside the kernel body these loops ran on spilled
// pointers (the body holds the whole schedule's state at 128 registers) — 10 % of the config-4 stall samples.
__device__ __noinline__ uint32_t x_stage(uint32_t *Lr, uint32_t *Sr, int sizeL, int offL, int sizeS, int offS, int lg, int sh,
                                         const uint32_t *idxL, const uint32_t *idxS, const uint32_t *cL, const uint32_t *gS,
                                         uint32_t rowb, uint32_t sbase, int lane, bool repl)
{
    uint32_t csum = 0;
    for (int idx = lane; idx < sizeL; idx += 32) {
        const int i = idx - offL;
        uint32_t d = 0;
        if (i == 0) d = 1;
        else if (i > 0 && i < lg) { d = idxL[i - 1]; csum += cL[d]; }
        Lr[idx] = d * rowb;
    }
    for (int idx = lane; idx < sizeS; idx += 32) {
        const int j = idx - offS;
        uint32_t d = 0;
        if (j == 0) d = 1;
        else if (j > 0 && j < sh) { d = idxS[j - 1]; csum += gS[d]; }
        // plain table: 2 bytes per column; replicated table (copy k in bank k): entry e at (e >> 1) * 128 + (e & 1) * 2
        Sr[idx] = (repl ? (((d >> 1) << 7) | ((d & 1u) << 1)) : d * 2u) + sbase;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(kFullM, csum, d);
    __syncwarp();
    return csum;
}

// ------------------------------------------------------------------------------------------------ batch dictionary
__device__ __forceinline__ uint32_t ov_any(int kind, const uint32_t *sg, int a, uint32_t x, uint32_t y, uint32_t &med)
{
    uint32_t c;
    if (kind == 0) { const uint32_t t = x & y; med = t ? t : (x | y); c = t ? 0u : 1u; }
    else if (kind == 1) ov_l1(a, x, y, med, c);
    else ov_explicit(sg, a, x, y, med, c);
    return c;
}

// Pass 1: every element of every string of the batch goes into a 64-slot global hash, through a per-CTA cache so that
// the global table sees each CTA's distinct elements once.  More than kDictMax distinct elements: overflow, the batch
// takes the per-pair path.
__global__ void __launch_bounds__(256) dict_scan_kernel(const MetricBatchDev b, uint32_t mask, DictDev *d)
{
    __shared__ uint32_t seen[256];
    __shared__ int s_stop;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) seen[i] = 0;
    if (threadIdx.x == 0) s_stop = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = warp; k < b.n_pairs; k += nw) {
        if (*reinterpret_cast<volatile int *>(&s_stop) || *reinterpret_cast<volatile uint32_t *>(&d->overflow)) break;
        const uint32_t n1 = b.len1[k], n2 = b.len2[k];
        const uint32_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
        for (uint32_t t = lane; t < n1 + n2; t += 32) {
            const uint32_t e = __ldg(t < n1 ? p1 + t : p2 + (t - n1)) & mask;
            if (!e) continue;
            uint32_t h = x_hash(e) & 255u, tries = 0;
            bool fresh = false;
            for (; tries < 256u; ++tries) {
                uint32_t old = seen[h];
                if (old == e) break;
                if (old == 0u) {
                    old = atomicCAS(&seen[h], 0u, e);
                    if (old == 0u) { fresh = true; break; }
                    if (old == e) break;
                }
                h = (h + 1u) & 255u;
            }
            if (tries >= 200u) { s_stop = 1; d->overflow = 1u; break; }
            if (!fresh) continue;
            uint32_t g = x_hash(e) & 63u;
            for (uint32_t tr = 0; tr < 64u; ++tr) {
                const uint32_t old = atomicCAS(&d->keys[g], 0u, e);
                if (old == e) break;
                if (old == 0u) {
                    if (atomicAdd(&d->count, 1u) >= (uint32_t)kDictMax) d->overflow = 1u;
                    break;
                }
                g = (g + 1u) & 63u;
                if (tr == 63u) d->overflow = 1u;
            }
        }
    }
}

// Pass 2 (one CTA): number the elements, indel costs, the 32 x 32 table in the folded form the fills consume.
__global__ void __launch_bounds__(1024) dict_build_kernel(DictDev *d, int kind, const uint32_t *sigma, int a, int dialect)
{
    __shared__ uint32_t s_el[32], s_cL[32], s_gS[32];
    __shared__ uint32_t s_D;
    __shared__ int s_min, s_max, s_bad, s_gapmed;
    const uint32_t gap = 1u << ((a - 1) & 31);
    if (threadIdx.x == 0) {
        uint32_t r = 0;
        for (int sl = 0; sl < 64; ++sl) {
            const uint32_t key = d->keys[sl];
            if (key && r < (uint32_t)kDictMax) { d->vals[sl] = 2u + r; s_el[2u + r] = key; ++r; }
        }
        s_el[0] = 0; s_el[1] = gap;
        s_D = r;
        s_min = -1; s_max = -1; s_bad = 0; s_gapmed = 0;
    }
    __syncthreads();
    const uint32_t D = s_D;
    if (threadIdx.x < 32) {
        const uint32_t dd = threadIdx.x;
        uint32_t md, c1 = 0, c2 = 0;
        if (dd >= 2u && dd < D + 2u) {
            c1 = ov_any(kind, sigma, a, s_el[dd], gap, md);         // x from the longer string against the gap
            c2 = ov_any(kind, sigma, a, gap, s_el[dd], md);
        }
        s_cL[dd] = c1; s_gS[dd] = c2;
        d->cL[dd] = c1; d->gS[dd] = c2;
        d->el[dd] = dd < D + 2u ? s_el[dd] : 0u;
    }
    __syncthreads();
    {
        const uint32_t dl = threadIdx.x >> 5, ds = threadIdx.x & 31u;
        int v;
        if (dl == 0u || ds == 0u || dl >= D + 2u || ds >= D + 2u) v = kBigSub;
        else if (dl == 1u || ds == 1u) v = -1;
        else {
            uint32_t md;
            const long long c = ov_any(kind, sigma, a, s_el[dl], s_el[ds], md);
            const long long f = 4ll * (c - (long long)s_cL[dl] - (long long)s_gS[ds]) - 1ll;
            if (md == gap) s_gapmed = 1;
            if (f < -32000ll || f >= (long long)kBigSub) s_bad = 1;
            v = (int)f;
            atomicMin(&s_min, v); atomicMax(&s_max, v);
        }
        d->tab[threadIdx.x] = (int16_t)v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        d->emin = s_min; d->emax = s_max;
        d->has_gap_median = (uint32_t)s_gapmed;
        // S2 relabels cells whose median is exactly the gap (+1 on the entry): rare, and per pair (only pairs that do
        // run the banded S2 schedule): such batches keep the per-pair tables
        d->ok = (!d->overflow && !s_bad && D >= 1u && !(s_gapmed && dialect == 2)) ? 1u : 0u;
    }
}

cudaError_t launch_dict(const ExplicitParams &p, DictDev *dict, int grid, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(dict, 0, sizeof(DictDev), stream);
    if (e != cudaSuccess) return e;
    const uint32_t mask = p.a >= 32 ? 0xffffffffu : ((1u << p.a) - 1u);
    dict_scan_kernel<<<grid * 4, 256, 0, stream>>>(p.b, mask, dict);
    dict_build_kernel<<<1, 1024, 0, stream>>>(dict, p.kind, p.sigma, p.a, p.dialect);
    return cudaGetLastError();
}

extern __shared__ __align__(16) char x_smem[];

template <bool WIDE>
__global__ void __launch_bounds__(512, 1) do_explicit_kernel(const ExplicitParams p)
{
    const MetricBatchDev &b = p.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int a = p.a;
    const int W = WIDE ? (a + 31) >> 5 : 1;                 // words per element
    // narrow: sigma[a*a] staged in shared memory; wide: sigma stays in global memory (L2), the first 64 bytes of
    // shared memory hold the gap element
    uint32_t *sg = WIDE ? const_cast<uint32_t *>(p.sigma) : reinterpret_cast<uint32_t *>(x_smem);
    uint32_t *gapw = reinterpret_cast<uint32_t *>(x_smem);
    const uint32_t sg_bytes = WIDE ? 64u : (((uint32_t)(a * a) * 4u + 15u) & ~15u);
    const uint32_t per_warp = 512u + p.seq_cap + 2u * p.smem_tab_entries;
    // batch-dictionary mode: [replicated table 64 KB | hash keys[64] vals[64] | cL[32] gS[32]] in front of the warps' windows
    // (the shared-memory layout depends on whether the host reserved room for the dictionary, p.dict != NULL; whether
    // the dictionary is USED is decided here, from what dict_build_kernel found)
    const bool dict_room = !WIDE && p.dict != nullptr;
    const bool dict = dict_room && p.dict->ok != 0u;
    char *dbase = x_smem + sg_bytes;
    uint32_t *d_keys = reinterpret_cast<uint32_t *>(dbase + 65536), *d_vals = d_keys + 64, *d_cL = d_vals + 64, *d_gS = d_cL + 32;
    uint32_t *s_meta = reinterpret_cast<uint32_t *>(x_smem + sg_bytes + (dict_room ? kDictSmemBytes : 0u) + (size_t)warp * per_warp);
    uint32_t *s_flags = s_meta, *s_nal = s_meta + 32, *s_off = s_meta + 64, *s_geo = s_meta + 96;
    uint32_t *seq = s_meta + 128;
    int16_t *stab = reinterpret_cast<int16_t *>(reinterpret_cast<char *>(seq) + p.seq_cap);
    if (WIDE) {
        if (threadIdx.x < kMaxW) gapw[threadIdx.x] = ((a - 1) >> 5) == (int)threadIdx.x ? 1u << ((a - 1) & 31) : 0u;
    } else if (p.kind == 2) {
        for (int i = threadIdx.x; i < a * a; i += blockDim.x) sg[i] = p.sigma[i];
    }
    if (dict) {
        // 32 copies of the 32 x 32 table, copy k wholly in bank k: word (e >> 1) * 32 + k holds entries e, e + 1
        uint32_t *rt = reinterpret_cast<uint32_t *>(dbase);
        const uint32_t *nat = reinterpret_cast<const uint32_t *>(p.dict->tab);
        for (uint32_t wi = threadIdx.x; wi < 512u * 32u; wi += blockDim.x) rt[wi] = __ldg(nat + (wi >> 5));
        for (uint32_t i = threadIdx.x; i < 64u; i += blockDim.x) { d_keys[i] = p.dict->keys[i]; d_vals[i] = p.dict->vals[i]; }
        for (uint32_t i = threadIdx.x; i < 32u; i += blockDim.x) { d_cL[i] = p.dict->cL[i]; d_gS[i] = p.dict->gS[i]; }
    }
    __syncthreads();
    const uint32_t rtab32 = (uint32_t)__cvta_generic_to_shared(dbase);
    const uint32_t lane_off = dict ? 4u * (uint32_t)lane : 0u;

    const uint32_t gap = WIDE ? 0u : 1u << ((a - 1) & 31);                      // narrow only
    const uint32_t mask = (WIDE || a >= 32) ? 0xffffffffu : ((1u << a) - 1u);
    const int G = p.pairs_per_warp;
    const size_t warp_global = (size_t)blockIdx.x * nwarps + warp;
    uint32_t *arena = p.arena + warp_global * (size_t)p.arena_words;
    uint32_t *ts = p.tscratch + warp_global * (size_t)p.tscratch_words;
    const bool want_out = b.out_med || b.out_a1 || b.out_a2 || b.out_len || b.out_ungapped;
    // optional multi-pass mode (PCGDO_EXPLICIT_MULTI_PASS): one Ukkonen offset per launch, pairs whose band must
    // double are queued for the next launch so that all warps run the same fill variant at the same time
    const uint32_t n_work = p.in_count ? *p.in_count : b.n_pairs;

    // Guided hand-out: a warp takes up to G pairs at a time (they are filled back to back and traced together, one lane
    // per pair), but never more than its share of half of what is left — the grabs shrink towards the end of the launch,
    // so that the warps finish together.  With fixed grabs of G the last wave left most of the SMs idle: a 65 536-pair
    // chunk of config 4 ran at 60 % of the full batch's rate.
    const uint32_t total_warps = gridDim.x * (uint32_t)nwarps;
    for (;;) {
        uint32_t base = 0, take = (uint32_t)G;
        if (lane == 0) {
            const uint32_t cur = *reinterpret_cast<volatile unsigned int *>(p.counter);
            const uint32_t rem = cur < n_work ? n_work - cur : 0u;
            take = min((uint32_t)G, max(1u, rem / (2u * total_warps)));
            base = atomicAdd(p.counter, take);
        }
        base = __shfl_sync(kFullM, base, 0);
        take = __shfl_sync(kFullM, take, 0);
        if (base >= n_work) break;
        const int npw = (int)min(take, n_work - base);
        int pi = 0;
        while (pi < npw) {
            const int start = pi;
            uint32_t cur = 0;
            for (; pi < npw; ++pi) {
                const uint32_t k = p.in_list ? p.in_list[base + pi] : base + pi;
                const uint32_t n1 = b.len1[k], n2 = b.len2[k];
                const uint32_t *p1 = b.elems + b.off1[k] * (uint64_t)W, *p2 = b.elems + b.off2[k] * (uint64_t)W;
                bool swapped = n1 > n2;                     // measureCharacters on the median strings
                if (n1 == n2) {
                    for (uint32_t bs = 0; bs < n1; bs += 32) {
                        const uint32_t i = bs + lane;
                        if constexpr (WIDE) {
                            const int c = i < n1 ? w_cmp(p1 + (size_t)i * W, p2 + (size_t)i * W, W) : 0;
                            const unsigned df = __ballot_sync(kFullM, c != 0);
                            if (df) { swapped = __shfl_sync(kFullM, c, __ffs(df) - 1) > 0; break; }
                        } else {
                            uint32_t x = 0, y = 0;
                            if (i < n1) { x = __ldg(p1 + i); y = __ldg(p2 + i); }
                            const unsigned df = __ballot_sync(kFullM, x != y);
                            if (df) {
                                const int src = __ffs(df) - 1;
                                swapped = __shfl_sync(kFullM, x, src) > __shfl_sync(kFullM, y, src);
                                break;
                            }
                        }
                    }
                }
                const uint32_t *pL = swapped ? p1 : p2, *pS = swapped ? p2 : p1;
                const int n = (int)(swapped ? n1 : n2), m = (int)(swapped ? n2 : n1);     // longer, lesser
                const int lg = n + 1, sh = m + 1;
                uint32_t gcnt = 0;
                if constexpr (WIDE) {
                    for (int i = lane; i < n; i += 32) gcnt += w_test(pL + (size_t)i * W, a - 1);
                    for (int j = lane; j < m; j += 32) gcnt += w_test(pS + (size_t)j * W, a - 1);
                } else {
                    for (int i = lane; i < n; i += 32) gcnt += (__ldg(pL + i) & gap) != 0u;
                    for (int j = lane; j < m; j += 32) gcnt += (__ldg(pS + j) & gap) != 0u;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) gcnt += __shfl_xor_sync(kFullM, gcnt, d);
                const long long Gc = gcnt, qd = n - m + 1, delta = p.delta;
                bool no_gain = m <= 4 || 2 * n >= 3 * m || delta == 0;
                if (p.dialect == 2) no_gain = no_gain || Gc >= n;
                const bool full = p.dialect == 0 || no_gain;
                const bool s2 = p.dialect == 2 && !no_gain;
                long long off = p.dialect == 2 ? 2 + Gc : 2;
                if (p.in_list) off = p.pair_off[k];                  // a later pass: the offset this pair was queued with
                bool requeue = false;
                long long jump_base = 0;                         // first skipped level while a jump of the schedule is unchecked
                bool may_jump = p.skip_ahead && !p.out_list;     // (one failed jump per pair, then level by level)

                // ---------------- number the distinct elements, build profiles, indel costs and the pair's table
                const uint32_t HL = x_hash_size((uint32_t)n), HS = x_hash_size((uint32_t)m);
                uint32_t *hkL = ts, *hvL = hkL + HL, *hkS = hvL + HL, *hvS = hkS + HS;
                uint32_t *elL = hvS + HS, *elS = elL + n, *idxL = elS + m, *idxS = idxL + n;
                const uint32_t *cL = idxS + m, *gS = cL + (n + 2);
                uint32_t *cLw = idxS + m, *gSw = cLw + (n + 2);
                uint32_t *mxL = gSw + (m + 2), *mxS = mxL + (size_t)n * a;
                uint32_t *tail = mxS + (size_t)m * a;
                bool ok = (size_t)(tail - ts) <= p.tscratch_words;
                uint32_t DL = 0, DS = 0, Wt = 0, entries = 0;
                bool gtable = false;
                const char *gtab = nullptr;
                uint32_t tab32 = 0;
                uint32_t bad = 0;
                int emin = -1, emax = -1;                   // range of the finite table entries (leading gap: -1)
                if (ok && dict) {
                    // table indices straight from the batch dictionary; no per-pair numbering, profiles or table
                    uint32_t miss = 0;
                    for (int t = lane; t < n + m; t += 32) {
                        const uint32_t e = __ldg(t < n ? pL + t : pS + (t - n)) & mask;
                        uint32_t v = 0;
                        if (e) {
                            uint32_t h = x_hash(e) & 63u, tries = 0;
                            while (d_keys[h] != e && tries < 64u) { h = (h + 1u) & 63u; ++tries; }
                            if (tries == 64u) miss = 1; else v = d_vals[h];
                        }
                        (t < n ? idxL : idxS)[t < n ? t : t - n] = v;
                    }
                    if (__any_sync(kFullM, miss != 0u)) ok = false;       // (cannot happen: the scan covered this batch)
                    cL = d_cL; gS = d_gS;
                    emin = p.dict->emin; emax = p.dict->emax;
                    tab32 = rtab32;
                    Wt = 1024u;                              // row stride of the replicated table: 2048 bytes
                    __syncwarp();
                } else if (ok) {
                    if constexpr (WIDE) {
                        DL = x_number_w(pL, n, W, HL, hkL, hvL, elL, idxL, lane);
                        DS = x_number_w(pS, m, W, HS, hkS, hvS, elS, idxS, lane);
                    } else {
                        DL = x_number(pL, n, mask, HL, hkL, hvL, elL, idxL, lane);
                        DS = x_number(pS, m, mask, HS, hkS, hvS, elS, idxS, lane);
                    }
                    Wt = DS + 2u;
                    entries = (DL + 2u) * Wt;
                    gtable = entries > p.smem_tab_entries;
                    if (gtable) ok = (size_t)(tail - ts) + (entries + 1u) / 2u <= p.tscratch_words && Wt * 2u * (DL + 2u) < 0x7fffffffu;
                }
                if (ok && !dict && p.kind != 2) {
                    // closed forms — discrete metric (Data/MetricRepresentation.hs:116-128): cost(x, y) = x & y ? 0 : 1;
                    // L1 norm (:166-190): interval logic — no profiles.  An element is elL[d] (narrow) or the string
                    // element at position elL[d] (wide).
                    auto ovc = [&](uint32_t dl, uint32_t ds, bool &is_gap) -> uint32_t {   // dl / ds: 1 = the gap element
                        if constexpr (WIDE) {
                            const uint32_t *x = dl == 1u ? gapw : pL + (size_t)elL[dl - 2u] * W;
                            const uint32_t *y = ds == 1u ? gapw : pS + (size_t)elS[ds - 2u] * W;
                            uint32_t md[kMaxW];
                            const uint32_t c = w_overlap(p.kind, sg, a, W, x, y, md);
                            is_gap = true;
                            for (int w = 0; w < W; ++w) is_gap = is_gap && md[w] == gapw[w];
                            return c;
                        } else {
                            const uint32_t x = dl == 1u ? gap : elL[dl - 2u], y = ds == 1u ? gap : elS[ds - 2u];
                            uint32_t md, c;
                            if (p.kind == 0) { const uint32_t t2 = x & y; md = t2 ? t2 : (x | y); c = t2 ? 0u : 1u; }
                            else ov_l1(a, x, y, md, c);
                            is_gap = md == gap;
                            return c;
                        }
                    };
                    bool ig;
                    for (uint32_t d = lane; d < DL + 2u; d += 32) cLw[d] = d >= 2u ? ovc(d, 1u, ig) : 0u;
                    for (uint32_t d = lane; d < DS + 2u; d += 32) gSw[d] = d >= 2u ? ovc(1u, d, ig) : 0u;
                    __syncwarp();
                    int16_t *tab = gtable ? reinterpret_cast<int16_t *>(tail) : stab;
                    for (uint32_t t = lane; t < entries; t += 32) {
                        const uint32_t dl = t / Wt, ds = t - dl * Wt;
                        int v;
                        if (dl == 0u || ds == 0u) v = kBigSub;
                        else if (dl == 1u || ds == 1u) v = -1;
                        else {
                            bool is_gap;
                            const long long c = ovc(dl, ds, is_gap);
                            const long long f = 4ll * (c - (long long)cL[dl] - (long long)gS[ds]) - 1ll + ((s2 && is_gap) ? 1ll : 0ll);
                            if (f < -32000ll || f >= (long long)kBigSub) bad = 1;
                            v = (int)f;
                            emin = min(emin, v); emax = max(emax, v);
                        }
                        tab[t] = (int16_t)v;
                    }
                    bad = __any_sync(kFullM, bad != 0u) ? 1u : 0u;
                    if (bad) ok = false;
                    gtab = reinterpret_cast<const char *>(tail);
                    tab32 = (uint32_t)__cvta_generic_to_shared(stab);
                    __syncwarp();
                } else if (ok && !dict) {
                    for (uint32_t t = lane; t < DL * (uint32_t)a; t += 32) {
                        const uint32_t d = t / a, s = t - d * a;
                        uint32_t mn = 0xffffffffu;
                        if constexpr (WIDE) mn = w_rowmin(sg + (size_t)s * a, pL + (size_t)elL[d] * W, W);
                        else for (uint32_t x = elL[d]; x; x &= x - 1) mn = min(mn, sg[s * a + __ffs(x) - 1]);
                        mxL[t] = mn;
                    }
                    for (uint32_t t = lane; t < DS * (uint32_t)a; t += 32) {
                        const uint32_t d = t / a, s = t - d * a;
                        uint32_t mn = 0xffffffffu;
                        if constexpr (WIDE) mn = w_rowmin(sg + (size_t)s * a, pS + (size_t)elS[d] * W, W);
                        else for (uint32_t x = elS[d]; x; x &= x - 1) mn = min(mn, sg[s * a + __ffs(x) - 1]);
                        mxS[t] = mn;
                    }
                    __syncwarp();
                    for (uint32_t d = lane; d < DL + 2u; d += 32) {
                        uint32_t c = 0;
                        if (d >= 2u) {
                            c = 0xffffffffu;
                            for (int s = 0; s < a; ++s) c = min(c, mxL[(d - 2u) * a + s] + sg[s * a + a - 1]);
                        }
                        cLw[d] = c;
                    }
                    for (uint32_t d = lane; d < DS + 2u; d += 32) {
                        uint32_t c = 0;
                        if (d >= 2u) {
                            c = 0xffffffffu;
                            for (int s = 0; s < a; ++s) c = min(c, mxS[(d - 2u) * a + s] + sg[s * a + a - 1]);
                        }
                        gSw[d] = c;
                    }
                    __syncwarp();
                    int16_t *tab = gtable ? reinterpret_cast<int16_t *>(tail) : stab;
                    for (uint32_t t = lane; t < entries; t += 32) {
                        const uint32_t dl = t / Wt, ds = t - dl * Wt;
                        int v;
                        if (dl == 0u || ds == 0u) v = kBigSub;
                        else if (dl == 1u || ds == 1u) v = -1;
                        else {
                            const uint32_t *ml = mxL + (size_t)(dl - 2u) * a, *ms = mxS + (size_t)(ds - 2u) * a;
                            uint32_t best = 0xffffffffu;
                            int nmin = 0, smin = -1;                 // median == {gap}  <=>  the only minimal symbol is a-1
                            for (int s = 0; s < a; ++s) {
                                const uint32_t c = ml[s] + ms[s];
                                if (c < best) { best = c; nmin = 1; smin = s; }
                                else if (c == best) ++nmin;
                            }
                            const long long f = 4ll * ((long long)best - (long long)cL[dl] - (long long)gS[ds]) - 1ll +
                                                ((s2 && nmin == 1 && smin == a - 1) ? 1ll : 0ll);
                            if (f < -32000ll || f >= (long long)kBigSub) bad = 1;
                            v = (int)f;
                            emin = min(emin, v); emax = max(emax, v);
                        }
                        tab[t] = (int16_t)v;
                    }
                    bad = __any_sync(kFullM, bad != 0u) ? 1u : 0u;
                    if (bad) ok = false;
                    gtab = reinterpret_cast<const char *>(tail);
                    tab32 = (uint32_t)__cvta_generic_to_shared(stab);
                    __syncwarp();
                }

                // packed 16-bit fill when the entries leave room in the 15-bit window (see do_linear.cu: rebase16)
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    emin = min(emin, __shfl_xor_sync(kFullM, emin, d));
                    emax = max(emax, __shfl_xor_sync(kFullM, emax, d));
                }
                const int s_minus = 1 - emin, s_plus = max(0, emax) + 2, spread16 = 0x7000 - 64 * s_minus - 64;
                bool packed = p.pack16 && 0x7000 + 64 * s_plus < 0x7800 && spread16 >= 8192;

                const MGeom gfull = make_mgeom(lg, sh, -(lg - 1), sh - 1);
                uint32_t fw = 0, csum = 0;
                MGeom g = gfull;
                int B = x_choose_b(gfull.wb, p.skip_b);
                uint32_t need = 0;
                int final_off = -1;
                bool used16 = false;
                while (ok) {
                    if (!full) {
                        const int a_off = (int)(off < m ? off : m);
                        g = make_mgeom(lg, sh, -(n - m) - a_off, a_off);
                        B = x_choose_b(g.wb, p.skip_b);
                        final_off = (int)(off > 0x7fffffff ? 0x7fffffff : off);
                    }
                    MSeq sl = {0, 0, 0, 0};
                    ok = B != 0;
                    if (ok) {
                        sl = m_seq_layout(g, B);
                        need = (m_dir_words(g, B) + m_ops_words(lg, sh) + 3u) & ~3u;
                        ok = 4u * (uint32_t)(sl.sizeL + sl.sizeS) <= p.seq_cap && need <= p.arena_words;
                    }
                    if (!ok) break;
                    if (cur + need > p.arena_words) break;
                    // stage row offsets (bytes into the table) and column offsets (+ the table's shared address)
                    uint32_t *Sr = seq, *Lr = seq + sl.sizeS;
                    csum = x_stage(Lr, Sr, sl.sizeL, sl.offL, sl.sizeS, sl.offS, lg, sh, idxL, idxS, cL, gS, Wt * 2u,
                                   gtable ? 0u : tab32, lane, dict);
                    uint32_t *slot = arena + cur;
                    used16 = packed && B >= 4;
                    if (used16) {
                        if (gtable) fw = x_dispatch16<true>(B, g, p.k_four, p.k_one, p.k_64k, gtab, Lr + sl.offL, Sr + sl.offS, slot, lane, (uint32_t)spread16, 0u);
                        else        fw = x_dispatch16<false>(B, g, p.k_four, p.k_one, p.k_64k, gtab, Lr + sl.offL, Sr + sl.offS, slot, lane, (uint32_t)spread16, lane_off);
                        if (fw == kOverflow16) { used16 = false; packed = false; }       // redo this and later fills in 32 bits
                    }
                    if (!used16) {
                        if (gtable) fw = x_dispatch<true>(B, g, p.k_four, p.k_minus1, p.k_one, gtab, Lr + sl.offL, Sr + sl.offS, slot, lane, 0u);
                        else        fw = x_dispatch<false>(B, g, p.k_four, p.k_minus1, p.k_one, gtab, Lr + sl.offL, Sr + sl.offS, slot, lane, lane_off);
                    }
                    __syncwarp();
                    if (full) break;
                    const long long cost = (long long)((fw >> 2) - (kBias4 >> 2) + csum);
                    if (p.dialect == 2) {
                        // The reference doubles the offset until the threshold exceeds the band's cost, re-filling at every
                        // level.  cost(offset) never grows with the offset, so a fill further up the schedule settles the
                        // levels it skipped: they would all have doubled if threshold(o) <= cost(later band) for the largest
                        // of them (thresholds grow with o).  After a level that doubles, `first_stop` is the first level that
                        // must stop given this level's cost (an upper bound on every later cost): the schedule jumps there
                        // when that level is at most skip_ahead (64) x the cells of this one, and falls back — to the first level that may
                        // stop given the jumped-to band's cost, sequentially from there — when the check fails.  Same final
                        // offset, band and alignment as level-by-level doubling (final_offset is part of what the parity tests compare).
                        auto thr2 = [&](long long o) { return (qd + o <= Gc) ? 0ll : delta * (qd + o - Gc); };
                        auto first_stop = [&](long long from, long long c) {
                            long long o = from;
                            while (!(qd + o > n) && !(thr2(o) > c)) o *= 2;
                            return o;
                        };
                        if (jump_base) {                       // this fill was a jump: were the skipped levels sure to double?
                            const long long base = jump_base;
                            jump_base = 0;
                            if (thr2(off / 2) > cost) { off = first_stop(base, cost); may_jump = false; continue; }
                        }
                        if (qd + off > n) break;
                        if (thr2(off) > cost) break;
                        const long long nxt = off * 2;
                        off = nxt;
                        if (may_jump) {
                            const long long L = first_stop(nxt, cost);
                            auto cells_of = [&](long long o) {
                                const long long a_o = o < m ? o : m, w = 2 * a_o + qd;
                                return (w < sh ? w : (long long)sh) * lg;
                            };
                            if (L > nxt && cells_of(L) <= (long long)p.skip_ahead * cells_of(nxt / 2)) { jump_base = nxt; off = L; }
                        }
                    } else {
                        long long thr = delta * (qd + off - Gc);
                        if (thr < 0) thr = 0;
                        if (thr <= cost) off *= 2; else break;
                    }
                    if (p.out_list) { requeue = true; break; }       // the doubled band is the next launch's work
                }
                if (ok && requeue) {
                    if (lane == 0) {
                        s_flags[pi] = 0;
                        p.pair_off[k] = off;
                        p.out_list[atomicAdd(p.out_count, 1u)] = k;
                    }
                    continue;
                }
                if (!ok) {
                    if (lane == 0) { s_flags[pi] = 0; atomicAdd(p.overflow_count, 1u); if (bad) atomicAdd(p.overflow_count + 1, 1u); }
                    continue;
                }
                if (cur + need > p.arena_words) break;       // flush what we have, then redo this pair
                if (lane == 0) {
                    s_flags[pi] = (swapped ? 1u : 0u) | 2u | (used16 ? 4u : 0u);
                    s_off[pi] = cur;
                    s_geo[pi] = (uint32_t)(uint16_t)(int16_t)g.dlo | ((uint32_t)(uint16_t)(int16_t)g.dhi << 16);
                    b.cost[k] = (int32_t)((fw >> 2) - (kBias4 >> 2) + csum);
                    if (b.final_offset) b.final_offset[k] = final_off;
                }
                if (b.cells) {
                    unsigned long long c = m_band_cells(g, lane, 32);
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(kFullM, c, d);
                    if (lane == 0) b.cells[k] = c;
                }
                cur += need;
                __syncwarp();
            }
            __syncwarp();
            if (!want_out) continue;
            const int pj = start + lane;
            if (pj < pi && (s_flags[pj] & 2u)) {
                const uint32_t k = p.in_list ? p.in_list[base + pj] : base + pj;
                const bool swapped = s_flags[pj] & 1u;
                const int n = (int)(swapped ? b.len1[k] : b.len2[k]), m = (int)(swapped ? b.len2[k] : b.len1[k]);
                const MGeom g = make_mgeom(n + 1, m + 1, (int)(int16_t)(s_geo[pj] & 0xffffu), (int)(int16_t)(s_geo[pj] >> 16));
                const int B = x_choose_b(g.wb, p.skip_b);
                uint32_t *slot = arena + s_off[pj];
                const uint32_t na = (s_flags[pj] & 4u) ? trace_lane16(slot, B, g.q0, g.lg, g.sh, slot + m_dir_words(g, B))
                                                       : m_trace(slot, 32u * B, g.q0, g.lg, g.sh, slot + m_dir_words(g, B));
                s_nal[pj] = na;
                if (b.out_len) b.out_len[k] = na;
            }
            __syncwarp();
            if (b.out_med || b.out_a1 || b.out_a2 || b.out_ungapped) {
                for (int pe = start; pe < pi; ++pe) {
                    if (!(s_flags[pe] & 2u)) continue;
                    const uint32_t k = p.in_list ? p.in_list[base + pe] : base + pe;
                    const bool swapped = s_flags[pe] & 1u;
                    const uint32_t *p1 = b.elems + b.off1[k] * (uint64_t)W, *p2 = b.elems + b.off2[k] * (uint64_t)W;
                    const int n = (int)(swapped ? b.len1[k] : b.len2[k]), m = (int)(swapped ? b.len2[k] : b.len1[k]);
                    const MGeom g = make_mgeom(n + 1, m + 1, (int)(int16_t)(s_geo[pe] & 0xffffu), (int)(int16_t)(s_geo[pe] >> 16));
                    const int B = x_choose_b(g.wb, p.skip_b);
                    const uint32_t *ops = arena + s_off[pe] + m_dir_words(g, B);
                    const uint64_t oo = (b.out_off ? b.out_off[k] : 0) * (uint64_t)W;
                    uint32_t *ou = b.out_ungapped ? b.out_ungapped + b.out_uoff[k] * (uint64_t)W : nullptr;
                    uint32_t nu;
                    if constexpr (WIDE)
                        nu = x_emit_w(sg, a, W, p.kind, swapped ? p1 : p2, swapped ? p2 : p1, gapw, ops, s_nal[pe], swapped,
                                      b.out_med ? b.out_med + oo : nullptr, b.out_a1 ? b.out_a1 + oo : nullptr,
                                      b.out_a2 ? b.out_a2 + oo : nullptr, lane, ou, b.ucap);
                    else
                        nu = x_emit(sg, a, p.kind, swapped ? p1 : p2, swapped ? p2 : p1, gap, mask, ops, s_nal[pe],
                                    swapped, b.out_med ? b.out_med + oo : nullptr, b.out_a1 ? b.out_a1 + oo : nullptr,
                                    b.out_a2 ? b.out_a2 + oo : nullptr, lane, ou, b.ucap);
                    if (b.out_ungapped && lane == 0) {
                        b.out_ulen[k] = nu < b.ucap ? nu : b.ucap;
                        if (nu > b.ucap) atomicAdd(p.overflow_count + 2, 1u);   // node string does not fit its slot
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------ dialect 3
// unboxedFullMatrixDO (DO/Pairwise/UnboxedFullMatrix.hs:48-144): the S2 cell over the full matrix, with its first column
// seeded from ROW 0 as the reference codes it — (i,0) = cost(left_i, gap) + m(0, i-1) (`ref (0, i - 1)`, :86-90).  Not on
// PCG's dispatch path (selectDynamicMetric never picks it); kept for the reference's own cross-implementation checks, so
// this kernel is plain: one warp per pair, lanes across an anti-diagonal, the overlap evaluated per cell (no memo table),
// values and one direction byte per cell in the warp's global scratch, traceback and medians by lane 0.
__global__ void __launch_bounds__(256) do_fullmatrix_kernel(const ExplicitParams p)
{
    const MetricBatchDev &b = p.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int a = p.a;
    __shared__ uint32_t sg[32 * 32];
    if (p.kind == 2)
        for (int i = threadIdx.x; i < a * a; i += blockDim.x) sg[i] = p.sigma[i];
    __syncthreads();
    const uint32_t gap = 1u << (a - 1), mask = a >= 32 ? 0xffffffffu : ((1u << a) - 1u);
    const size_t warp_global = (size_t)blockIdx.x * nwarps + warp, warps_total = (size_t)gridDim.x * nwarps;
    uint32_t *ts = p.tscratch + warp_global * (size_t)p.tscratch_words;
    for (size_t k = warp_global; k < b.n_pairs; k += warps_total) {
        const uint32_t n1 = b.len1[k], n2 = b.len2[k];
        const uint32_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
        bool swapped = n1 > n2;                                     // measureCharacters on the median strings
        if (n1 == n2) {
            for (uint32_t bs = 0; bs < n1; bs += 32) {
                const uint32_t i = bs + lane;
                uint32_t x = 0, y = 0;
                if (i < n1) { x = __ldg(p1 + i); y = __ldg(p2 + i); }
                const unsigned df = __ballot_sync(kFullM, x != y);
                if (df) {
                    const int src = __ffs(df) - 1;
                    swapped = __shfl_sync(kFullM, x, src) > __shfl_sync(kFullM, y, src);
                    break;
                }
            }
        }
        const uint32_t *top = swapped ? p1 : p2, *left = swapped ? p2 : p1;        // longer = columns, lesser = rows
        const int n = (int)(swapped ? n1 : n2), m = (int)(swapped ? n2 : n1);
        const size_t Wd = (size_t)n + 1;
        uint32_t *va = ts, *vb = va + (m + 2), *vc = vb + (m + 2), *row0 = vc + (m + 2);
        uint8_t *dir = reinterpret_cast<uint8_t *>(row0 + Wd);
        if ((size_t)3 * (m + 2) + Wd + ((size_t)(m + 1) * Wd + 3) / 4 > p.tscratch_words) {
            if (lane == 0) atomicAdd(p.overflow_count, 1u);
            continue;
        }
        for (int d = 0; d <= m + n; ++d) {
            const int ilo = max(0, d - n), ihi = min(m, d);
            for (int i = ilo + lane; i <= ihi; i += 32) {
                const int j = d - i;
                uint32_t v, md;
                uint8_t dr;
                if (i == 0 && j == 0) { v = 0; dr = 0; }
                else if (i == 0) { v = vb[0] + ov_any(p.kind, sg, a, gap, __ldg(top + j - 1) & mask, md); dr = 1; }
                else if (j == 0) { v = row0[i - 1] + ov_any(p.kind, sg, a, __ldg(left + i - 1) & mask, gap, md); dr = 2; }
                else {
                    const uint32_t te = __ldg(top + j - 1) & mask, le = __ldg(left + i - 1) & mask;
                    uint32_t mdd, tmp;
                    const uint32_t cd = va[i - 1] + ov_any(p.kind, sg, a, te, le, mdd);
                    const uint32_t cl = vb[i] + ov_any(p.kind, sg, a, te, gap, tmp);
                    const uint32_t cu = vb[i - 1] + ov_any(p.kind, sg, a, gap, le, tmp);
                    v = cd; dr = 0;                                  // minimum of (cost, direction): Diag < Left < Up
                    if (cl < v) { v = cl; dr = 1; }
                    if (cu < v) { v = cu; dr = 2; }
                    if (dr == 0 && mdd == gap) dr = 1;               // getMinimalResult
                }
                vc[i] = v;
                if (i == 0) row0[j] = v;
                dir[(size_t)i * Wd + j] = dr;
            }
            __syncwarp();
            uint32_t *t = va; va = vb; vb = vc; vc = t;
        }
        // traceback + outputs (Pairwise/Internal.hs:531-581), lane 0; vb holds the last anti-diagonal
        if (lane == 0) {
            b.cost[k] = (int32_t)vb[m];
            if (b.cells) b.cells[k] = (unsigned long long)(m + 1) * (unsigned long long)(n + 1);
            if (b.final_offset) b.final_offset[k] = -1;
            uint32_t nal = 0;
            for (int i = m, j = n; i > 0 || j > 0; ++nal) {
                const uint8_t dr = dir[(size_t)i * Wd + j];
                i -= (dr != 1);
                j -= (dr != 2);
            }
            if (b.out_len) b.out_len[k] = nal;
            const uint64_t oo = b.out_off ? b.out_off[k] : 0;
            uint32_t c = nal;
            for (int i = m, j = n; i > 0 || j > 0;) {
                const uint8_t dr = dir[(size_t)i * Wd + j];
                const uint32_t x = dr != 1 ? (__ldg(left + i - 1) & mask) : 0u, y = dr != 2 ? (__ldg(top + j - 1) & mask) : 0u;
                uint32_t md;
                if (dr == 1) ov_any(p.kind, sg, a, gap, y, md);
                else if (dr == 2) ov_any(p.kind, sg, a, x, gap, md);
                else ov_any(p.kind, sg, a, x, y, md);
                --c;
                if (b.out_med) b.out_med[oo + c] = md;
                if (b.out_a1) b.out_a1[oo + c] = swapped ? y : x;
                if (b.out_a2) b.out_a2[oo + c] = swapped ? x : y;
                i -= (dr != 1);
                j -= (dr != 2);
            }
        }
        __syncwarp();
    }
}

cudaError_t launch_fullmatrix(const ExplicitParams &p, int grid, cudaStream_t stream)
{
    do_fullmatrix_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// per-warp global scratch (32-bit words) for strings of up to max_len elements over `a` symbols:
// two hash tables, distinct elements, indices, indel costs, profiles, and the table itself (capped)
size_t explicit_scratch_words(int max_len, int a, size_t table_cap_entries)
{
    const size_t n = (size_t)max_len;
    const size_t H = x_hash_size((uint32_t)max_len);
    size_t entries = (n + 2) * (n + 2);
    if (entries > table_cap_entries) entries = table_cap_entries;
    return 4 * H + 4 * n + 2 * (n + 2) + 2 * n * (size_t)a + (entries + 1) / 2 + 16;
}

cudaError_t launch_explicit(const ExplicitParams &p, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(do_explicit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(do_explicit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        once.mark(dev);
    }
    if (p.a > 32) do_explicit_kernel<true><<<grid, block, smem_bytes, stream>>>(p);
    else          do_explicit_kernel<false><<<grid, block, smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace pcgdo
