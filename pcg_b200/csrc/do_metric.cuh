// do_metric.cuh — pieces shared by the large-alphabet DO kernels (do_metric.cu: discrete metric in registers,
// do_explicit.cu: explicit symbol-change matrices through a per-pair memo table).
#pragma once
#include "do_device.cuh"

namespace pcgdo {

constexpr unsigned kFullM = 0xffffffffu;

struct MetricBatchDev {
    uint32_t        n_pairs;
    const uint32_t *elems;
    const uint64_t *off1, *off2;      // element offsets
    const uint32_t *len1, *len2;
    const uint64_t *out_off;          // element offsets into the three output arrays
    uint32_t       *out_med, *out_a1, *out_a2;
    uint32_t       *out_len;
    int32_t        *cost;
    uint64_t       *cells;
    int32_t        *final_offset;     // may be NULL
    // post-order mode (do_explicit.cu only): the median with its gap elements dropped, at element offset
    // out_uoff[k], at most ucap elements, length to out_ulen
    uint32_t       *out_ungapped;
    const uint64_t *out_uoff;
    uint32_t       *out_ulen;
    uint32_t        ucap;
};

struct MetricParams {
    MetricBatchDev b;
    int       a;             // alphabet size (incl. gap), <= 24
    int       dialect;       // 0 S1 full, 1 S1 Ukkonen, 2 S2 Ukkonen
    uint32_t *arena;
    uint32_t  arena_words;
    uint32_t  seq_cap;       // bytes of staging per warp
    int       pairs_per_warp;
    unsigned int *counter;
    unsigned int *overflow_count;
    uint32_t  k_four, k_minus1;
};

// band as a diagonal range of MY matrix (rows = longer incl. leading gap, columns = lesser incl. leading gap)
struct MGeom { int lg, sh, dlo, dhi, pad, q0, wb, T; };

__host__ __device__ inline MGeom make_mgeom(int lg, int sh, int dlo, int dhi)
{
    MGeom g;
    g.lg = lg; g.sh = sh;
    if (dlo < -(lg - 1)) dlo = -(lg - 1);
    if (dhi > sh - 1) dhi = sh - 1;
    g.dlo = dlo; g.dhi = dhi;
    g.pad = 1 + ((1 - dlo) & 1);
    g.q0 = -dlo + g.pad;
    g.wb = dhi - dlo + 1 + g.pad;
    const int q_last = (sh - 1) - (lg - 1) + g.q0;
    g.T = ((lg - 1) + (sh - 1) - (q_last & 1)) / 2 + 1;
    return g;
}
__host__ __device__ inline int m_choose_b(int wb)
{
    if (wb + 1 <= 64) return 2;
    if (wb + 1 <= 128) return 4;
    if (wb + 1 <= 256) return 8;
    if (wb + 1 <= 512) return 16;
    if (wb + 1 <= 1024) return 32;
    return 0;
}
__host__ __device__ inline uint32_t m_dir_words(const MGeom &g, int B) { return (uint32_t)((g.T + 15) / 16) * 32u * (uint32_t)B; }
__host__ __device__ inline uint32_t m_ops_words(int lg, int sh) { return (uint32_t)(lg + sh) / 16u + 2u; }
struct MSeq { int offL, sizeL, offS, sizeS; };
__host__ __device__ inline MSeq m_seq_layout(const MGeom &g, int B)
{
    const int H = B / 2, tpad = (g.T + 15) & ~15;
    MSeq s;
    s.offL = 32 * H;
    int needL = tpad + g.q0 / 2 + 2;
    s.sizeL = s.offL + (needL > g.lg ? needL : g.lg) + 2;
    s.offS = g.q0 / 2;
    int needS = tpad - g.q0 / 2 + 32 * H + 2;
    if (needS < g.sh) needS = g.sh;
    s.sizeS = s.offS + needS + 2;
    return s;
}
__host__ __device__ inline unsigned long long m_band_cells(const MGeom &g, int i0, int step)
{
    unsigned long long c = 0;
    for (int i = i0; i < g.lg; i += step) {
        int lo = i + g.dlo; if (lo < 0) lo = 0;
        int hi = i + g.dhi; if (hi > g.sh - 1) hi = g.sh - 1;
        if (hi >= lo) c += (unsigned long long)(hi - lo + 1);
    }
    return c;
}

__device__ __forceinline__ uint32_t m_imad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int B>
__device__ __forceinline__ void m_store(uint32_t *dst, const uint32_t (&dw)[B], int n)
{
    const int sh = 2 * (16 - n);
    const uint32_t ones = 0x55555555u >> sh;
    uint32_t o[B];
#pragma unroll
    for (int m = 0; m < B; ++m) o[m] = (dw[m] + ones) << sh;
    if constexpr (B >= 4) {
#pragma unroll
        for (int m = 0; m < B; m += 4) *reinterpret_cast<uint4 *>(dst + m) = make_uint4(o[m], o[m + 1], o[m + 2], o[m + 3]);
    } else {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1]);
    }
}

// codes: 0 Diag, 1 consume longer (Haskell LeftArrow), 2 consume lesser (UpArrow)
__device__ __forceinline__ uint32_t m_trace(const uint32_t *dirw, uint32_t wq, int q0, int lg, int sh, uint32_t *ops)
{
    int s = (lg - 1) + (sh - 1), q = (sh - 1) - (lg - 1) + q0;        // (i + j, diagonal): see trace_lane16
    uint32_t n = 0, acc = 0, w = 0;
    int sft = 32;
    while (s > 0) {
        if (sft == 32) {
            const int t = (s - (q & 1)) >> 1;
            w = __ldcg(dirw + ((uint32_t)(t >> 4) * wq + (uint32_t)q));
            sft = 2 * (15 - (t & 15));
        }
        const uint32_t code = (w >> sft) & 3u;
        acc |= code << (2 * (n & 15));
        ++n;
        if ((n & 15) == 0) { ops[(n >> 4) - 1] = acc; acc = 0; }
        if (code == 0u) { s -= 2; sft += 2; }
        else { s -= 1; q += (code == 1u) ? 1 : -1; sft = 32; }
    }
    if (n & 15) ops[n >> 4] = acc;
    return n;
}

}  // namespace pcgdo
