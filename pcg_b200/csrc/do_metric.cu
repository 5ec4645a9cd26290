// do_metric.cu — the Haskell DO dialects for alphabets too large for a dense table (a > 8), sm_100a.
//
// Reference paths (DO = lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization):
//   S1 full matrix   naiveDO / naiveDOMemo / unboxedSwappingDO   DO/Pairwise/Internal.hs:389-424,
//                                                                DO/Pairwise/UnboxedSwapping.hs:58-124
//   S1 Ukkonen       ukkonenDO                                   DO/Pairwise/Ukkonen/Internal.hs:79-202, Ribbon.hs:100-155
//   S2 Ukkonen       unboxedUkkonenFullSpaceDO (production, a>8) DO/Pairwise/UnboxedUkkonenFullSpace.hs:59-276, 641-815
// with the discrete metric's closed-form overlap (lib/core/tcm/src/Data/MetricRepresentation.hs:116-128)
// computed per cell in registers — no tcm-memo hash lookups (lib/tcm-memo), no per-cell FFI.
// PARITY UNPINNED upstream for these paths (no executable Haskell here, no golden vector): the checker is the
// CPU restatement kept with the tests.  The Ukkonen variants re-fill the band from scratch at each offset
// of the reference's doubling schedule (the reference expands incrementally; same cells, same rules).
//
// Same machinery as do_linear.cu: one warp per pair, lanes own B consecutive diagonals of the band, two
// anti-diagonals per iteration, cell value 4*(v - C(i) - G(j) + BIAS) | priority so that one unsigned 3-way
// min gives cost and direction with the reference's tie order Diag > Left (consume longer) > Up (consume
// lesser); S2's "median == gap => relabel Diag as Left" is folded into the diagonal candidate's low bit.
// Elements are 32-bit words (a <= 24 on this path), packed in shared memory as  x | (x has no gap bit) << 24
// so that the folded-out indel costs of a cell are one add and one shift away.
#include "do_metric.cuh"

namespace pcgdo {

// 4*(sigma(x,y) - c(x) - g(y)) - 1, plus 1 when S2 and the median is exactly the gap (Diag relabelled Left)
template <bool S2>
__device__ __forceinline__ uint32_t m_sub(uint32_t xv, uint32_t yv, uint32_t gap)
{
    const uint32_t t = xv & yv & 0xffffffu;
    const uint32_t cg = (xv + yv) >> 24;                 // c(x) + g(y) in {0, 1, 2}
    int e = (t == 0u) ? 3 : -1;
    if (S2) e = (t == gap) ? 0 : e;
    return (uint32_t)(e - 4 * (int)cg);
}

template <int B, bool S2, bool TAIL>
__device__ __forceinline__ void m_chunk(uint32_t (&w)[B], const uint32_t (&penc)[B], uint32_t (&dw)[B],
                                        uint32_t (&lw)[B / 2], uint32_t (&sw)[B / 2], uint32_t k4, uint32_t km1,
                                        uint32_t gap, const uint32_t *&lptr, const uint32_t *&sptr, int n_iter)
{
    constexpr int H = B / 2;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (TAIL && r >= n_iter) break;
        const int tt = r % H;
        const uint32_t wl = __shfl_up_sync(kFullM, w[B - 1], 1);
#pragma unroll
        for (int u = 0; u < H; ++u) {
            const int m = 2 * u;
            const uint32_t left = (u == 0) ? wl : w[m > 0 ? m - 1 : 0];
            const uint32_t sub = m_sub<S2>(lw[(tt - u + H) % H], sw[(tt + u) % H], gap);
            const uint32_t t1 = __viaddmin_u32(w[m], sub, w[m + 1]);
            const uint32_t mn = __viaddmin_u32(left, 1u, t1);
            const uint32_t wn = (mn & ~3u) | penc[m];
            dw[m] = m_imad(wn, km1, m_imad(dw[m], k4, mn));
            w[m] = wn;
        }
        sw[tt] = sptr[r];
        const uint32_t wr = __shfl_down_sync(kFullM, w[0], 1);
#pragma unroll
        for (int u = 0; u < H; ++u) {
            const int m = 2 * u + 1;
            const uint32_t up = (u == H - 1) ? wr : w[m + 1 < B ? m + 1 : B - 1];
            const uint32_t sub = m_sub<S2>(lw[(tt - u + H) % H], sw[(tt + u + 1) % H], gap);
            const uint32_t t1 = __viaddmin_u32(w[m], sub, up);
            const uint32_t mn = __viaddmin_u32(w[m - 1], 1u, t1);
            const uint32_t wn = (mn & ~3u) | penc[m];
            dw[m] = m_imad(wn, km1, m_imad(dw[m], k4, mn));
            w[m] = wn;
        }
        lw[(tt + 1) % H] = lptr[r];
    }
    lptr += 16;
    sptr += 16;
}


template <int B, bool S2>
__device__ __noinline__ uint32_t m_fill(const MGeom g, uint32_t k4, uint32_t km1, uint32_t gap, const uint32_t *Lb,
                                        const uint32_t *Sb, uint32_t *slot, int lane)
{
    constexpr int H = B / 2;
    uint32_t w[B], penc[B], dw[B], lw[H], sw[H];
    const int qb = lane * B;
    const uint32_t sub_gg = m_sub<S2>(gap, gap, gap);        // both leading gaps: c = g = 0
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int q = qb + m;
        penc[m] = 1u | ((q < g.pad || q >= g.wb) ? 0x80000000u : 0u);
        asm volatile("" : "+r"(penc[m]));
        w[m] = (q == g.q0) ? (kBias4 - sub_gg) : kInfW;
        dw[m] = 0;
    }
    const int I0 = g.q0 / 2 - lane * H, J0 = -(g.q0 / 2) + lane * H;
#pragma unroll
    for (int u = 0; u < H; ++u) {
        lw[(H - u) % H] = Lb[I0 - u];
        sw[u] = Sb[J0 + u];
    }
    const uint32_t *lptr = Lb + (I0 + 1), *sptr = Sb + (J0 + H);
    uint32_t *out = slot + qb;
    const int nfull = g.T >> 4, rem = g.T & 15;
    for (int c = 0; c < nfull; ++c) {
        m_chunk<B, S2, false>(w, penc, dw, lw, sw, k4, km1, gap, lptr, sptr, 16);
        m_store<B>(out, dw, 16);
        out += 32 * B;
    }
    if (rem) {
#pragma unroll
        for (int m = 0; m < B; ++m) dw[m] = 0;
        m_chunk<B, S2, true>(w, penc, dw, lw, sw, k4, km1, gap, lptr, sptr, rem);
        m_store<B>(out, dw, rem);
    }
    const int q_last = (g.sh - 1) - (g.lg - 1) + g.q0, m_last = q_last % B;
    uint32_t v = 0;
#pragma unroll
    for (int m = 0; m < B; ++m)
        if (m == m_last) v = w[m];
    return __shfl_sync(kFullM, v, q_last / B);
}

template <bool S2>
__device__ __forceinline__ uint32_t m_dispatch(int B, const MGeom &g, uint32_t k4, uint32_t km1, uint32_t gap,
                                               const uint32_t *Lb, const uint32_t *Sb, uint32_t *slot, int lane)
{
    switch (B) {
        case 2:  return m_fill<2, S2>(g, k4, km1, gap, Lb, Sb, slot, lane);
        case 4:  return m_fill<4, S2>(g, k4, km1, gap, Lb, Sb, slot, lane);
        case 8:  return m_fill<8, S2>(g, k4, km1, gap, Lb, Sb, slot, lane);
        case 16: return m_fill<16, S2>(g, k4, km1, gap, Lb, Sb, slot, lane);
        default: return m_fill<32, S2>(g, k4, km1, gap, Lb, Sb, slot, lane);
    }
}


__device__ __forceinline__ uint32_t m_median(uint32_t x, uint32_t y) { return (x & y) ? (x & y) : (x | y); }

// pL / pS: longer / lesser strings (no leading gap). swapped: in1 is the longer one.
__device__ __forceinline__ void m_emit(const uint32_t *pL, const uint32_t *pS, uint32_t gap, const uint32_t *ops,
                                       uint32_t nal, bool swapped, uint32_t *om, uint32_t *o1, uint32_t *o2, int lane)
{
    uint32_t base_i = 0, base_j = 0;
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
        if (c < nal) {
            const uint32_t y = code < 2u ? __ldg(pL + il) : 0u;
            const uint32_t x = (code == 0u || code == 2u) ? __ldg(pS + js) : 0u;
            const uint32_t md = m_median(x ? x : gap, y ? y : gap);
            if (om) om[c] = md;
            if (o1) o1[c] = swapped ? y : x;
            if (o2) o2[c] = swapped ? x : y;
        }
        base_i += __popc(bl);
        base_j += __popc(bs);
    }
}

extern __shared__ __align__(16) char m_smem[];

__global__ void __launch_bounds__(512, 1) do_metric_kernel(const MetricParams p)
{
    const MetricBatchDev &b = p.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t per_warp = 512u + p.seq_cap;
    uint32_t *s_meta = reinterpret_cast<uint32_t *>(m_smem + (size_t)warp * per_warp);
    uint32_t *s_flags = s_meta, *s_nal = s_meta + 32, *s_off = s_meta + 64, *s_geo = s_meta + 96;   // geo: dlo | dhi packed
    uint32_t *seq = s_meta + 128;
    const uint32_t gap = 1u << (p.a - 1);
    const int G = p.pairs_per_warp;
    uint32_t *arena = p.arena + ((size_t)blockIdx.x * nwarps + warp) * (size_t)p.arena_words;
    const bool want_out = b.out_med || b.out_a1 || b.out_a2 || b.out_len;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(p.counter, (unsigned)G);
        base = __shfl_sync(kFullM, base, 0);
        if (base >= b.n_pairs) break;
        const int npw = (int)min((uint32_t)G, b.n_pairs - base);
        int pi = 0;
        while (pi < npw) {
            const int start = pi;
            uint32_t cur = 0;
            for (; pi < npw; ++pi) {
                const uint32_t k = base + pi;
                const uint32_t n1 = b.len1[k], n2 = b.len2[k];
                const uint32_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
                // measureCharacters on the median strings: shorter first, then lexicographic; equal -> no swap
                bool swapped = n1 > n2;
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
                const uint32_t *pL = swapped ? p1 : p2, *pS = swapped ? p2 : p1;
                const int n = (int)(swapped ? n1 : n2), m = (int)(swapped ? n2 : n1);     // longer, lesser
                const int lg = n + 1, sh = m + 1;
                // G = elements containing the gap bit; Delta = 1 for the discrete metric
                uint32_t gcnt = 0;
                for (int i = lane; i < n; i += 32) gcnt += (__ldg(pL + i) & gap) != 0u;
                for (int j = lane; j < m; j += 32) gcnt += (__ldg(pS + j) & gap) != 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) gcnt += __shfl_xor_sync(kFullM, gcnt, d);
                const long long Gc = gcnt, qd = n - m + 1;
                bool no_gain = m <= 4 || 2 * n >= 3 * m;
                if (p.dialect == 2) no_gain = no_gain || Gc >= n;
                const bool full = p.dialect == 0 || no_gain;
                const bool s2 = p.dialect == 2 && !no_gain;
                long long off = p.dialect == 2 ? 2 + Gc : 2;
                // worst-case scratch: the band can grow to the full matrix
                const MGeom gfull = make_mgeom(lg, sh, -(lg - 1), sh - 1);
                const int Bfull = m_choose_b(gfull.wb);
                bool ok = true;
                uint32_t fw = 0, csum = 0;
                MGeom g = gfull;
                int B = Bfull;
                uint32_t need = 0;
                int final_off = -1;
                for (;;) {
                    if (!full) {
                        const int a_off = (int)(off < m ? off : m);
                        g = make_mgeom(lg, sh, -(n - m) - a_off, a_off);
                        B = m_choose_b(g.wb);
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
                    // stage: x | (no gap bit) << 24; the leading elements are the gap itself
                    uint32_t *Sr = seq, *Lr = seq + sl.sizeS;
                    csum = 0;
                    for (int idx = lane; idx < sl.sizeL; idx += 32) {
                        const int i = idx - sl.offL;
                        uint32_t v = 0;
                        if (i == 0) v = gap;
                        else if (i > 0 && i < lg) { const uint32_t e = __ldg(pL + i - 1) & 0xffffffu; const uint32_t c = (e & gap) ? 0u : 1u; v = e | (c << 24); csum += c; }
                        Lr[idx] = v;
                    }
                    for (int idx = lane; idx < sl.sizeS; idx += 32) {
                        const int j = idx - sl.offS;
                        uint32_t v = 0;
                        if (j == 0) v = gap;
                        else if (j > 0 && j < sh) { const uint32_t e = __ldg(pS + j - 1) & 0xffffffu; const uint32_t c = (e & gap) ? 0u : 1u; v = e | (c << 24); csum += c; }
                        Sr[idx] = v;
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(kFullM, csum, d);
                    __syncwarp();
                    uint32_t *slot = arena + cur;
                    if (s2) fw = m_dispatch<true>(B, g, p.k_four, p.k_minus1, gap, Lr + sl.offL, Sr + sl.offS, slot, lane);
                    else    fw = m_dispatch<false>(B, g, p.k_four, p.k_minus1, gap, Lr + sl.offL, Sr + sl.offS, slot, lane);
                    __syncwarp();
                    if (full) break;
                    const long long cost = (long long)((fw >> 2) - (kBias4 >> 2) + csum);
                    if (p.dialect == 2) {
                        if (qd + off > n) break;
                        const long long thr = (qd + off <= Gc) ? 0 : (qd + off - Gc);
                        if (thr <= cost) off *= 2; else break;
                    } else {
                        long long thr = qd + off - Gc;
                        if (thr < 0) thr = 0;
                        if (thr <= cost) off *= 2; else break;
                    }
                }
                if (!ok) {
                    if (lane == 0) { s_flags[pi] = 0; atomicAdd(p.overflow_count, 1u); }
                    continue;
                }
                if (cur + need > p.arena_words) break;       // flush what we have, then redo this pair
                if (lane == 0) {
                    s_flags[pi] = (swapped ? 1u : 0u) | 2u;
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
                const uint32_t k = base + pj;
                const bool swapped = s_flags[pj] & 1u;
                const int n = (int)(swapped ? b.len1[k] : b.len2[k]), m = (int)(swapped ? b.len2[k] : b.len1[k]);
                const MGeom g = make_mgeom(n + 1, m + 1, (int)(int16_t)(s_geo[pj] & 0xffffu), (int)(int16_t)(s_geo[pj] >> 16));
                const int B = m_choose_b(g.wb);
                uint32_t *slot = arena + s_off[pj];
                const uint32_t na = m_trace(slot, 32u * B, g.q0, g.lg, g.sh, slot + m_dir_words(g, B));
                s_nal[pj] = na;
                if (b.out_len) b.out_len[k] = na;
            }
            __syncwarp();
            if (b.out_med || b.out_a1 || b.out_a2) {
                for (int pe = start; pe < pi; ++pe) {
                    if (!(s_flags[pe] & 2u)) continue;
                    const uint32_t k = base + pe;
                    const bool swapped = s_flags[pe] & 1u;
                    const uint32_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
                    const int n = (int)(swapped ? b.len1[k] : b.len2[k]), m = (int)(swapped ? b.len2[k] : b.len1[k]);
                    const MGeom g = make_mgeom(n + 1, m + 1, (int)(int16_t)(s_geo[pe] & 0xffffu), (int)(int16_t)(s_geo[pe] >> 16));
                    const int B = m_choose_b(g.wb);
                    const uint32_t *ops = arena + s_off[pe] + m_dir_words(g, B);
                    const uint64_t oo = b.out_off[k];
                    m_emit(swapped ? p1 : p2, swapped ? p2 : p1, gap, ops, s_nal[pe], swapped,
                           b.out_med ? b.out_med + oo : nullptr, b.out_a1 ? b.out_a1 + oo : nullptr,
                           b.out_a2 ? b.out_a2 + oo : nullptr, lane);
                }
            }
            __syncwarp();
        }
    }
}

uint32_t metric_seq_cap(int max_len)
{
    // both strings as 32-bit words plus the window paddings of the widest (B = 32) variant
    return 4u * (uint32_t)(2 * max_len + 2200);
}
uint32_t metric_words_for(uint32_t n1, uint32_t n2)
{
    const int lg = (int)(n1 > n2 ? n1 : n2) + 1, sh = (int)(n1 > n2 ? n2 : n1) + 1;
    const MGeom g = make_mgeom(lg, sh, -(lg - 1), sh - 1);
    const int B = m_choose_b(g.wb);
    return B ? ((m_dir_words(g, B) + m_ops_words(lg, sh) + 3u) & ~3u) : 0u;
}

cudaError_t launch_metric(const MetricParams &p, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(do_metric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        once.mark(dev);
    }
    do_metric_kernel<<<grid, block, smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace pcgdo
