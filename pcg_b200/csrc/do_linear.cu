// do_linear.cu — linear-gap pairwise direct optimization ("C-linear" dialect) for sm_100a.
//
// Replaces, bit for bit, the reference's align2d pipeline
//   EXT/c_alignment_interface.c:16-177  (long/short choice, band parameter)
//   EXT/alignCharacters.c:388-658, 675-1374 (banded row fill), :4404-4507 (traceback),
//   :4569-4594 (gapped median)            [EXT = lib/core/ffi/external-direct-optimization]
// with one persistent kernel.  One WARP owns one pair at a time:
//
//   fill   lanes statically own B consecutive diagonals (q = j - i + q0) of the reference's band;
//          the wavefront advances two anti-diagonals per iteration (even q, then odd q), each cell
//          updated in place in registers, neighbours across lanes via one __shfl per half step.
//          Cell values are stored as 4*(v - C(i) - G(j) + BIAS) | priority-code, so that
//            up   needs no add, left only the +1 priority bump, diag one table add,
//          and a single unsigned 3-way min yields both the cost and the reference's traceback
//          choice (ALIGN > DELETE > INSERT).  The 2-bit choices accumulate per diagonal as
//          D = 4*D + mn - w (two IMADs: the FMA pipe is otherwise idle while the ALU pipe is the
//          bottleneck) and are flushed to HBM every 16 iterations as B consecutive words per lane
//          (coalesced 128-bit stores): word[(t >> 4) * Wq + q], cell t at bits 2*(15 - (t & 15)).
//   trace  after a warp has filled up to 32 pairs it walks all of them at once, one lane per pair,
//          reading direction words with 16 consecutive ALIGN steps per word, and writes a packed
//          2-bit op stream.
//   emit   warp-cooperative per pair: prefix sums over the op stream give the element indices, the
//          median comes from the dense table, and the three aligned strings leave as coalesced
//          32-bit stores.
//
// No tensor cores: the recurrence is an integer min-plus DP.
#include "do_device.cuh"
#include "do_pack16.cuh"

namespace pcgdo {


__device__ __forceinline__ uint32_t ld_u16(const char *p) { return *reinterpret_cast<const uint16_t *>(p); }
__device__ __forceinline__ uint32_t ld_u8(const char *p) { return *reinterpret_cast<const uint8_t *>(p); }
// a*b + c on the FMA pipe; b is a run-time value so ptxas cannot strength-reduce it onto the ALU pipe
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// the TCM table is written once before the block-wide barrier and never again: a plain (non-volatile) load
__device__ __forceinline__ uint32_t lds_s16(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
struct FillConst { uint32_t k4, km1, lstride, k64k; };

// Smallest supported diagonals-per-lane for a band of `wb` diagonals (0 = does not fit a warp).
__host__ __device__ inline int choose_b(int wb)
{
    // + 1: the last diagonal of the warp (q = 32*B - 1) stays masked, like q = 0, so that the
    // block-edge shuffles need no special case in lanes 0 and 31
    if (wb + 1 <= 64)   return 2;
    if (wb + 1 <= 128)  return 4;
    if (wb + 1 <= 256)  return 8;
    if (wb + 1 <= 512)  return 16;
    if (wb + 1 <= 1024) return 32;
    return 0;
}

// Staging layout of one pair in the warp's shared-memory window: the shorter string S as u16 entries
// already in "table address" form, then the longer string L as raw bytes (its table row offset is one
// IMAD away), each with zero padding on both sides so that every lane's sliding window may run off the
// ends of the strings.  Sizes are in elements.
struct SeqLayout { int offL, sizeL, offS, sizeS; };

__host__ __device__ inline SeqLayout seq_layout(const Geom &g, int B)
{
    const int H = B / 2;
    const int tpad = (g.T + 15) & ~15;
    SeqLayout s;
    s.offL = 32 * H;
    int needL = tpad + g.q0 / 2 + 2;
    s.sizeL = s.offL + (needL > g.lg ? needL : g.lg) + 2;
    s.offS = g.q0 / 2;
    int needS = tpad - g.q0 / 2 + 32 * H + 2;
    if (needS < g.sh) needS = g.sh;
    s.sizeS = s.offS + needS + 2;
    s.sizeL = (s.sizeL + 3) & ~3;
    s.sizeS = (s.sizeS + 1) & ~1;
    return s;
}
__host__ __device__ inline uint32_t seq_bytes(const SeqLayout &s) { return 2u * (uint32_t)s.sizeS + (uint32_t)s.sizeL; }

__host__ __device__ inline uint32_t dir_words(const Geom &g, int B) { return (uint32_t)((g.T + 15) / 16) * 32u * (uint32_t)B; }
__host__ __device__ inline uint32_t ops_words(const Geom &g) { return (uint32_t)(g.lg + g.sh) / 16u + 2u; }

// ---------------------------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------------------------
template <int B, bool TAIL>
__device__ __forceinline__ void run_chunk(uint32_t (&w)[B], const uint32_t (&penc)[B], uint32_t (&dw)[B],
                                          uint32_t (&lw)[B / 2], uint32_t (&sw)[B / 2], const FillConst fc,
                                          uint32_t tab_lane, const char *&lptr, const char *&sptr, int n_iter)
{
    constexpr int H = B / 2;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        if (TAIL && r >= n_iter) break;
        const int tt = r % H;
        // ---- even diagonals: anti-diagonal s = 2t.  Lane 0 receives its own value here: it only feeds the
        // always-masked diagonal q = 0.
        const uint32_t wl = __shfl_up_sync(kFull, w[B - 1], 1);
#pragma unroll
        for (int u = 0; u < H; ++u) {
            const int m = 2 * u;
            const uint32_t left = (u == 0) ? wl : w[m > 0 ? m - 1 : 0];
            const uint32_t up = w[m + 1];
            const uint32_t sub = lds_s16(imad(lw[(tt - u + H) % H], fc.lstride, sw[(tt + u) % H]));
            const uint32_t t1 = __viaddmin_u32(w[m], sub, up);
            const uint32_t mn = __viaddmin_u32(left, 1u, t1);
            const uint32_t wn = (mn & ~3u) | penc[m];
            dw[m] = imad(wn, fc.km1, imad(dw[m], fc.k4, mn));
            w[m] = wn;
        }
        sw[tt] = ld_u16(sptr + 2 * r) + tab_lane;
        // ---- odd diagonals: anti-diagonal s = 2t + 1 (lane 31's own value only feeds masked q = 32B-1)
        const uint32_t wr = __shfl_down_sync(kFull, w[0], 1);
#pragma unroll
        for (int u = 0; u < H; ++u) {
            const int m = 2 * u + 1;
            const uint32_t left = w[m - 1];
            const uint32_t up = (u == H - 1) ? wr : w[m + 1 < B ? m + 1 : B - 1];
            const uint32_t sub = lds_s16(imad(lw[(tt - u + H) % H], fc.lstride, sw[(tt + u + 1) % H]));
            const uint32_t t1 = __viaddmin_u32(w[m], sub, up);
            const uint32_t mn = __viaddmin_u32(left, 1u, t1);
            const uint32_t wn = (mn & ~3u) | penc[m];
            dw[m] = imad(wn, fc.km1, imad(dw[m], fc.k4, mn));
            w[m] = wn;
        }
        lw[(tt + 1) % H] = ld_u8(lptr + r);
    }
    lptr += 16;
    sptr += 32;
}

// dw holds sum_k (code_k - 1) * 4^(n-1-k) over the n cells of the chunk: add the ones back, left-align
template <int B>
__device__ __forceinline__ void store_words(uint32_t *dst, const uint32_t (&dw)[B], int n)
{
    const int sh = 2 * (16 - n);
    const uint32_t ones = 0x55555555u >> sh;
    uint32_t o[B];
#pragma unroll
    for (int m = 0; m < B; ++m) o[m] = (dw[m] + ones) << sh;
    if constexpr (B >= 4) {
#pragma unroll
        for (int m = 0; m < B; m += 4)
            *reinterpret_cast<uint4 *>(dst + m) = make_uint4(o[m], o[m + 1], o[m + 2], o[m + 3]);
    } else {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1]);
    }
}

// Lb / Sb point at the entries of index 0 (the leading gap) of the staged strings; tab_lane is the
// shared-window address of this lane's copy of the TCM table.
template <int B>
__device__ __noinline__ uint32_t fill_pair(const Geom g, const FillConst fc, uint32_t tab_lane,
                                           const char *Lb, const char *Sb, uint32_t *slot, int lane, int sub4gg)
{
    constexpr int H = B / 2;
    uint32_t w[B], penc[B], dw[B], lw[H], sw[H];
    const int qb = lane * B;
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int q = qb + m;
        penc[m] = 1u | ((q < g.pad || q >= g.wb) ? 0x80000000u : 0u);
        asm volatile("" : "+r"(penc[m]));        // keep it in a register: ptxas otherwise re-derives it per cell
        w[m] = (q == g.q0) ? (kBias4 - (uint32_t)sub4gg) : kInfW;
        dw[m] = 0;
    }
    const int I0 = g.q0 / 2 - lane * H;      // row handled by this lane's first even diagonal at t = 0
    const int J0 = -(g.q0 / 2) + lane * H;   // its column
#pragma unroll
    for (int u = 0; u < H; ++u) {
        lw[(H - u) % H] = ld_u8(Lb + (I0 - u));
        sw[u] = ld_u16(Sb + 2 * (J0 + u)) + tab_lane;
    }
    const char *lptr = Lb + (I0 + 1);
    const char *sptr = Sb + 2 * (J0 + H);
    uint32_t *out = slot + qb;
    const int nfull = g.T >> 4, rem = g.T & 15;
    for (int c = 0; c < nfull; ++c) {
        run_chunk<B, false>(w, penc, dw, lw, sw, fc, tab_lane, lptr, sptr, 16);
        store_words<B>(out, dw, 16);
        out += 32 * B;
    }
    if (rem) {
#pragma unroll
        for (int m = 0; m < B; ++m) dw[m] = 0;
        run_chunk<B, true>(w, penc, dw, lw, sw, fc, tab_lane, lptr, sptr, rem);
        store_words<B>(out, dw, rem);
    }
    const int q_last = (g.sh - 1) - (g.lg - 1) + g.q0;
    const int m_last = q_last % B;
    uint32_t v = 0;
#pragma unroll
    for (int m = 0; m < B; ++m)
        if (m == m_last) v = w[m];
    return __shfl_sync(kFull, v, q_last / B);
}


// ---------------------------------------------------------------------------------------------
// packed fill: two diagonals per 32-bit register, VIADDMNMX.U16x2
//
// Register R[m] (m < H = B/2) holds diagonal m of the lane in its low half and diagonal m + H in its high half
// (same parity, H even), each as a 16-bit cell value 4*(v - C(i) - G(j)) + base | code with bit 15 = unreachable.
// One packed update = two cells: two table fetches merged by one IMAD, two VIADDMNMX.U16x2, one LOP3 to
// re-normalise, one LOP3 + one IMAD to append both 2-bit choices to a packed direction word — 10 instructions
// for 2 cells instead of 14.  Neighbours of a packed pair are the adjacent registers; only the lane edges need
// a shuffle plus a funnel shift.  The 15-bit window follows the values: every 32 iterations the warp's finite
// maximum is moved back to kTop16 (rebase16); if the finite spread of the sweep ever exceeds what the window
// holds, the pair is redone by the 32-bit fill (never silently wrong).
// Direction words: word[(t >> 3) * 32*H + lane*H + m], cell t of diagonal m (m + H) at bits 2*(7 - (t & 7))
// of the low (high) half.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// (a & b) | c as ONE LOP3: ptxas splits the C expression in two for about half of the updates
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// The H packed updates of an iteration (H / 2 even-diagonal registers, then H / 2 odd ones) are numbered f = 0 .. H-1; the
// two table fetches of update f + D are issued right before the arithmetic of update f, into a ring of D register pairs
// (they depend on the string windows only), so that a fetch has D updates of arithmetic between issue and use.
template <int B> __host__ __device__ constexpr int fetch_dist() { return B >= 32 ? 4 : (B >= 8 ? 2 : 1); }

// fetches of update g in the frame of iteration tt (g >= H: update g - H of the next iteration)
template <int B>
__device__ __forceinline__ void fetch16(int tt, int g, const uint32_t (&lw)[B / 2], const uint32_t (&sw)[B / 2], const FillConst fc,
                                        uint32_t &lo, uint32_t &hi)
{
    constexpr int H = B / 2, Q = H / 2, U = H / 2;
    const int t2 = tt + g / H, gg = g % H, odd = gg >= U ? 1 : 0, u = gg - U * odd;
    lo = lds_u16(imad(lw[(t2 - u + 2 * H) % H], fc.lstride, sw[(t2 + u + odd) % H]));
    hi = lds_u16(imad(lw[(t2 - u - Q + 2 * H) % H], fc.lstride, sw[(t2 + u + Q + odd) % H]));
}

template <int B, bool TAIL>
__device__ __forceinline__ void run_chunk16(uint32_t (&R)[B / 2], const uint32_t (&penc)[B / 2], uint32_t (&dw)[B / 2],
                                            uint32_t (&lw)[B / 2], uint32_t (&sw)[B / 2], uint32_t (&p_lo)[fetch_dist<B>()],
                                            uint32_t (&p_hi)[fetch_dist<B>()], const FillConst fc,
                                            uint32_t tab_lane, const char *&lptr, const char *&sptr, int n_iter,
                                            uint32_t *&out)
{
    constexpr int H = B / 2, U = H / 2, CH = pack16_chunk(B), D = fetch_dist<B>();
    static_assert(D <= U, "fetch distance");
#pragma unroll
    for (int r = 0; r < CH; ++r) {
        if (TAIL && r >= n_iter) break;
        const int tt = r % H;
        const uint32_t s_new = ld_u16(sptr + 2 * r) + tab_lane;
        const uint32_t l_new = ld_u8(lptr + r);
        uint32_t edge = 0;
#pragma unroll
        for (int f = 0; f < H; ++f) {
            if (f + D == U) sw[tt] = s_new;                 // before the first odd-diagonal fetch of this iteration
            if (f + D == H) lw[(tt + 1) % H] = l_new;       // before the first fetch of the next one
            const uint32_t sub2 = imad(p_hi[f % D], fc.k64k, p_lo[f % D]);
            fetch16<B>(tt, f + D, lw, sw, fc, p_lo[(f + D) % D], p_hi[(f + D) % D]);
            uint32_t mn;
            int m;
            if (f < U) {            // even diagonals (registers with even m): anti-diagonal s = 2t
                m = 2 * f;
                if (f == 0) {
                    const uint32_t prev = __shfl_up_sync(kFull, R[H - 1], 1);
                    edge = __funnelshift_l(prev, R[H - 1], 16);      // (own diag H-1) << 16 | (left lane's diag B-1)
                }
                const uint32_t left = (m == 0) ? edge : R[m > 0 ? m - 1 : 0];
                const uint32_t t1 = __viaddmin_u16x2(R[m], sub2, R[m + 1 < H ? m + 1 : H - 1]);
                mn = __viaddmin_u16x2(left, 0x00010001u, t1);
            } else {                // odd diagonals: anti-diagonal s = 2t + 1
                m = 2 * (f - U) + 1;
                if (f == U) {
                    const uint32_t next = __shfl_down_sync(kFull, R[0], 1);
                    edge = __funnelshift_r(R[0], next, 16);          // (right lane's diag 0) << 16 | (own diag H)
                }
                const uint32_t up = (m == H - 1) ? edge : R[m + 1 < H ? m + 1 : H - 1];
                const uint32_t t1 = __viaddmin_u16x2(R[m], sub2, up);
                mn = __viaddmin_u16x2(R[m - 1 >= 0 ? m - 1 : 0], 0x00010001u, t1);
            }
            R[m] = and_or(mn, 0xfffcfffcu, penc[m]);
            const uint32_t c2 = mn & 0x00030003u;
            dw[m] = (r % 8 == 0) ? c2 : imad(dw[m], fc.k4, c2);
        }
        if (r % 8 == 7) {
            store_words16<H>(out, dw, 8);
            out += 32 * H;
        } else if (TAIL && r == n_iter - 1) {
            store_words16<H>(out, dw, (r % 8) + 1);
            out += 32 * H;
        }
    }
    lptr += CH;
    sptr += 2 * CH;
    if constexpr (CH % H != 0) { rot16<H, CH % H>(lw); rot16<H, CH % H>(sw); }
}

// Returns the final cell's 4*(v - C - G) + kBias4 | code in the 32-bit convention of fill_pair, or kOverflow16.
template <int B>
__device__ __noinline__ uint32_t fill_pair16(const Geom g, const FillConst fc, uint32_t tab_lane, const char *Lb,
                                             const char *Sb, uint32_t *slot, int lane, int sub4gg, uint32_t spread_limit)
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
            const uint32_t w1 = (q == g.q0) ? ((kTop16 - (uint32_t)sub4gg) & 0xffffu) : 0x8001u;
            pe |= p1 << (16 * h);
            wv |= w1 << (16 * h);
        }
        penc[m] = pe;
        asm volatile("" : "+r"(penc[m]));
        R[m] = wv;
        dw[m] = 0;
    }
    const int I0 = g.q0 / 2 - lane * H;
    const int J0 = -(g.q0 / 2) + lane * H;
#pragma unroll
    for (int u = 0; u < H; ++u) {
        lw[(H - u) % H] = ld_u8(Lb + (I0 - u));
        sw[u] = ld_u16(Sb + 2 * (J0 + u)) + tab_lane;
    }
    const char *lptr = Lb + (I0 + 1);
    const char *sptr = Sb + 2 * (J0 + H);
    uint32_t *out = slot + lane * H;
    uint32_t p_lo[fetch_dist<B>()], p_hi[fetch_dist<B>()];
#pragma unroll
    for (int g0 = 0; g0 < fetch_dist<B>(); ++g0) fetch16<B>(0, g0, lw, sw, fc, p_lo[g0], p_hi[g0]);
    constexpr int CH = pack16_chunk(B), RB = 32 / CH;        // re-base every 32 iterations
    const int nfull = g.T / CH, rem = g.T % CH;
    int total_shift = 0;
    bool ok = true;
    for (int c = 0; c < nfull; ++c) {
        if (c && c % RB == 0) ok = rebase16<H>(R, total_shift, spread_limit) && ok;
        run_chunk16<B, false>(R, penc, dw, lw, sw, p_lo, p_hi, fc, tab_lane, lptr, sptr, CH, out);
    }
    if (rem) {
        if (nfull && nfull % RB == 0) ok = rebase16<H>(R, total_shift, spread_limit) && ok;
        run_chunk16<B, true>(R, penc, dw, lw, sw, p_lo, p_hi, fc, tab_lane, lptr, sptr, rem, out);
    }
    // one last look at the window: everything since the previous check stayed inside it iff that check passed
    ok = rebase16<H>(R, total_shift, spread_limit) && ok;
    if (!ok) return kOverflow16;
    const int q_last = (g.sh - 1) - (g.lg - 1) + g.q0;
    const int m_last = q_last % B;
    uint32_t v = 0;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        if (m == m_last) v = R[m] & 0xffffu;
        if (m + H == m_last) v = R[m] >> 16;
    }
    v = __shfl_sync(kFull, v, q_last / B);
    const int x4 = (int)(v & 0x7ffcu) - (int)kTop16 - total_shift;      // 4 * (v - C - G)
    return (uint32_t)((int)kBias4 + x4) | 1u;
}


// ---------------------------------------------------------------------------------------------
// multi-pair packed fill: FOUR pairs per warp, 8 lanes x 28 diagonals each
//
// A band of 129 .. 223 diagonals (1 kb strings: 200) leaves 22 % of a 32 x 8 warp as padding, and the per-iteration
// overhead of the wavefront (two shuffles, the funnel shifts, the window loads, the loop) is paid for 4 packed updates.
// Here a warp sweeps four pairs at once: lanes 8s .. 8s+7 own the diagonals of pair s, 28 per lane = 14 packed registers,
// so the same overhead is spread over 14 updates and 224 diagonal slots serve a pair instead of 256.  Everything that
// depends on the pair (geometry, string windows, masks) is per-lane data already; the shuffles stay warp-wide because
// what crosses a segment border only ever lands in the always-masked diagonals q = 0 and q = 223 of a pair.
// Pairs of a group may differ in length: all four run max(T) iterations.  Past a pair's own last cell the strings read
// as padding (table entry kBigSub), moves along rows / columns are free in the folded cell values, and so the final
// cell's value is carried unchanged down its diagonal — it is read after the common last iteration.
// Direction words: word[(t >> 3) * 448 + lane * 14 + m] — the diagonals of a pair are consecutive words (two diagonals 14
// apart share a word), so the trace's next word is usually in the 32-byte sector it already holds.
// ---------------------------------------------------------------------------------------------
// Shapes built: LP = 8 lanes x 28 diagonals (H = 14 registers; 4 pairs per warp; bands of 169 .. 223 diagonals: ~1 kb
// strings), LP = 6 x 28 (5 pairs per warp, two lanes idle; 141 .. 167) and LP = 5 x 28 (6 pairs, two lanes idle; 129 .. 139):
// the 300 .. 500-element medians of tree scoring, and LP = 4 lanes x 32 diagonals (H = 16; 8 pairs per warp; bands of
// 65 .. 127: its ~250-element leaves; off by default, see do_api.cu).  Direction words per 8 iterations of one group: 32 * H.
template <int H> __host__ __device__ inline uint32_t multi_dir_words(int tmax) { return (uint32_t)((tmax + 7) / 8) * 32u * (uint32_t)H; }

// Staging of one pair of a group: both strings as raw bytes (element 0 = padding), zero-padded for `tmax` iterations.
template <int LP, int H>
__host__ __device__ inline SeqLayout seq_layout_multi(const Geom &g, int tmax)
{
    const int tpad = (tmax + 15) & ~15;
    SeqLayout s;
    s.offL = LP * H;
    int needL = tpad + g.q0 / 2 + 2;
    s.sizeL = s.offL + (needL > g.lg ? needL : g.lg) + 2;
    s.offS = g.q0 / 2;
    int needS = tpad - g.q0 / 2 + LP * H + 2;
    if (needS < g.sh) needS = g.sh;
    s.sizeS = s.offS + needS + 2;
    s.sizeL = (s.sizeL + 3) & ~3;
    s.sizeS = (s.sizeS + 3) & ~3;
    return s;
}

// table offset of a lesser-string element (the dense kernels stage it precomputed as u16; here it is one LOP3 + IMAD
// per lane and iteration)
__device__ __forceinline__ uint32_t s_entry(uint32_t e, bool repl) { return repl ? (((e >> 1) << 7) | ((e & 1u) << 1)) : (e << 1); }

// Fetch distance of the multi-pair fill (see fetch16 / run_chunk16: the fetches of update f + D are issued before the
// arithmetic of update f).  With 4 warps per scheduler nothing else covers the shared-memory latency; half an iteration
// ahead measured as good as a whole batch up front.
template <int H> __host__ __device__ constexpr int multi_dist() { return H >= 16 ? 4 : (H / 2 < 7 ? H / 2 : 7); }   // (H = 16: the register file)

// s_entry without the branch: e * c1 - (e & 1) * c2 with (c1, c2) = (64, 62) for the replicated table, (2, 0) otherwise
__device__ __forceinline__ uint32_t s_entry2(uint32_t e, uint32_t c1, uint32_t c2neg, uint32_t base)
{
    return imad(e & 1u, c2neg, imad(e, c1, base));
}

template <int LP, int H, bool TAIL>
__device__ __forceinline__ void run_chunk16q(uint32_t (&R)[H], const uint32_t (&penc)[H], uint32_t (&dw)[H],
                                             uint32_t (&lw)[H], uint32_t (&sw)[H], uint32_t (&p_lo)[multi_dist<H>()],
                                             uint32_t (&p_hi)[multi_dist<H>()], const FillConst fc, uint32_t tab_lane, uint32_t sc1, uint32_t sc2,
                                             const char *&lptr, const char *&sptr, int n_iter, uint32_t *&out, int n_st)
{
    constexpr int U = H / 2, CH = 8, D = multi_dist<H>();
    static_assert(D >= 1 && D <= U && H > CH && H % 2 == 0, "shape");
#pragma unroll
    for (int r = 0; r < CH; ++r) {
        if (TAIL && r >= n_iter) break;
        const int tt = r;
        const uint32_t s_new = s_entry2(ld_u8(sptr + r), sc1, sc2, tab_lane);
        const uint32_t l_new = ld_u8(lptr + r);
        uint32_t edge = 0;
#pragma unroll
        for (int f = 0; f < H; ++f) {
            // the window slots are replaced once nothing issued later still wants the old element: the lesser string's
            // before the first odd-diagonal fetch of this iteration, the longer string's before the first fetch of the next
            if (f + D == U) sw[tt] = s_new;
            if (f + D == H) lw[(tt + 1) % H] = l_new;
            const uint32_t sub2 = imad(p_hi[f % D], fc.k64k, p_lo[f % D]);
            fetch16<2 * H>(tt, f + D, lw, sw, fc, p_lo[(f + D) % D], p_hi[(f + D) % D]);
            if (f < U) {            // ---- even diagonals (registers with even m): anti-diagonal s = 2t
                const int m = 2 * f;
                if (f == 0) {
                    const uint32_t prev = __shfl_up_sync(kFull, R[H - 1], 1);
                    edge = __funnelshift_l(prev, R[H - 1], 16);      // (own diag H-1) << 16 | (left lane's diag B-1)
                }
                const uint32_t left = (m == 0) ? edge : R[m > 0 ? m - 1 : 0];
                const uint32_t t1 = __viaddmin_u16x2(R[m], sub2, R[m + 1 < H ? m + 1 : H - 1]);
                const uint32_t mn = __viaddmin_u16x2(left, 0x00010001u, t1);
                R[m] = and_or(mn, 0xfffcfffcu, penc[m]);
                const uint32_t c2 = mn & 0x00030003u;
                dw[m] = (r == 0) ? c2 : imad(dw[m], fc.k4, c2);
            } else {                // ---- odd diagonals: anti-diagonal s = 2t + 1
                const int m = 2 * (f - U) + 1;
                if (f == U) {
                    const uint32_t next = __shfl_down_sync(kFull, R[0], 1);
                    edge = __funnelshift_r(R[0], next, 16);          // (right lane's diag 0) << 16 | (own diag H)
                }
                const uint32_t up = (m == H - 1) ? edge : R[m + 1 < H ? m + 1 : H - 1];
                const uint32_t t1 = __viaddmin_u16x2(R[m], sub2, up);
                const uint32_t mn = __viaddmin_u16x2(R[m - 1 >= 0 ? m - 1 : 0], 0x00010001u, t1);
                R[m] = and_or(mn, 0xfffcfffcu, penc[m]);
                const uint32_t c2 = mn & 0x00030003u;
                dw[m] = (r == 0) ? c2 : imad(dw[m], fc.k4, c2);
            }
        }
        if (r == CH - 1 || (TAIL && r == n_iter - 1)) {
            const int sh = 2 * (CH - 1 - r);          // left-align a partial group
            // (All registers, also those beyond the band: storing only the band's left partially written 32-byte sectors
            // behind — lanes are 56 bytes apart — which L2 completes by READING them back: DRAM reads 26.7 -> 33.8 GB per
            // launch for 2 GB fewer written.  n_st = 0 still skips a lane that is wholly beyond the band and 64-byte aligned.)
#pragma unroll
            for (int m = 0; m < H; m += 2)
                if (n_st) *reinterpret_cast<uint2 *>(out + m) = make_uint2(dw[m] << sh, dw[m + 1] << sh);
            out += 32 * H;
        }
    }
    lptr += CH;
    sptr += CH;
    rot16<H, CH % H>(lw);
    rot16<H, CH % H>(sw);
}

// Reductions over the LP lanes of one pair (sbase = its first lane).  LP need not be a power of two (5 lanes x 28
// diagonals serve six pairs per warp, lanes 30 / 31 idle along): then the values are gathered one by one.
template <int LP, class F>
__device__ __forceinline__ uint32_t seg_reduce(uint32_t v, int sbase, F f)
{
    if constexpr ((LP & (LP - 1)) == 0) {
#pragma unroll
        for (int d = LP / 2; d > 0; d >>= 1) v = f(v, __shfl_xor_sync(kFull, v, d));
        return v;
    } else {
        uint32_t r = __shfl_sync(kFull, v, sbase);
#pragma unroll
        for (int i = 1; i < LP; ++i) r = f(r, __shfl_sync(kFull, v, sbase + i));
        return r;
    }
}

// rebase16 over the LP lanes of one pair
template <int LP, int H>
__device__ __forceinline__ bool rebase16q(uint32_t (&R)[H], int &total_shift, uint32_t spread_limit, int sbase)
{
    uint32_t mx = 0u, mnv = 0xffffffffu;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        mx = __vmaxu2(mx, R[m] ^ 0x80008000u);
        mnv = __vminu2(mnv, R[m]);
    }
    uint32_t hi = max(mx & 0xffffu, mx >> 16), lo = min(mnv & 0xffffu, mnv >> 16);
    hi = seg_reduce<LP>(hi, sbase, [](uint32_t x, uint32_t y) { return max(x, y); });
    lo = seg_reduce<LP>(lo, sbase, [](uint32_t x, uint32_t y) { return min(x, y); });
    if (hi < 0x8000u) return true;
    const uint32_t maxf = hi & 0x7ffcu, minf = lo & 0xfffcu;
    const int delta = (int)kTop16 - (int)maxf;
    const uint32_t d2 = ((uint32_t)delta & 0xffffu) * 0x00010001u;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        const uint32_t unreach = ((R[m] >> 15) & 0x00010001u) * 0xffffu;
        R[m] = __vadd2(R[m], d2 & ~unreach);
    }
    total_shift += delta;
    return maxf - minf <= spread_limit;
}

// g, Lb, Sb: of the lane's own pair.  tmax: iterations of the group (warp-uniform).  slot: the group's direction words.
template <int LP, int H>
__device__ __noinline__ uint32_t fill_multi16(const Geom g, int tmax, const FillConst fc, uint32_t tab_lane, bool repl,
                                             const char *Lb, const char *Sb, uint32_t *slot, int lane, int sub4gg,
                                             uint32_t spread_limit)
{
    constexpr int B = 2 * H, D = multi_dist<H>();
    uint32_t R[H], penc[H], dw[H], lw[H], sw[H];
    const int ln = lane % LP, sbase = lane - ln;      // (idle lanes behind the last pair: a phantom pair, results unused)
    const int qb = ln * B;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        uint32_t pe = 0, wv = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = qb + m + h * H;
            const uint32_t p1 = 1u | ((q < g.pad || q >= g.wb) ? 0x8000u : 0u);
            const uint32_t w1 = (q == g.q0) ? ((kTop16 - (uint32_t)sub4gg) & 0xffffu) : 0x8001u;
            pe |= p1 << (16 * h);
            wv |= w1 << (16 * h);
        }
        penc[m] = pe;
        asm volatile("" : "+r"(penc[m]));
        R[m] = wv;
        dw[m] = 0;
    }
    const int I0 = g.q0 / 2 - ln * H;
    const int J0 = -(g.q0 / 2) + ln * H;
#pragma unroll
    for (int u = 0; u < H; ++u) {
        lw[(H - u) % H] = ld_u8(Lb + (I0 - u));
        sw[u] = s_entry(ld_u8(Sb + (J0 + u)), repl) + tab_lane;
    }
    const char *lptr = Lb + (I0 + 1);
    const char *sptr = Sb + (J0 + H);
    uint32_t *out = slot + lane * H;
    // a lane wholly beyond the band stores nothing when that leaves no 32-byte sector half written (H = 16: 64 bytes per lane)
    const int n_st = (H % 8 == 0 && qb >= g.wb) ? 0 : 1;
    const uint32_t sc1 = repl ? 64u : 2u, sc2 = repl ? (uint32_t)-62 : 0u;
    uint32_t p_lo[D], p_hi[D];
#pragma unroll
    for (int g0 = 0; g0 < D; ++g0) fetch16<B>(0, g0, lw, sw, fc, p_lo[g0], p_hi[g0]);
    const int nfull = tmax / 8, rem = tmax % 8;
    int total_shift = 0;
    bool ok = true;
    for (int c = 0; c < nfull; ++c) {
        if (c && c % 4 == 0) ok = rebase16q<LP, H>(R, total_shift, spread_limit, sbase) && ok;       // every 32 iterations
        run_chunk16q<LP, H, false>(R, penc, dw, lw, sw, p_lo, p_hi, fc, tab_lane, sc1, sc2, lptr, sptr, 8, out, n_st);
    }
    if (rem) {
        if (nfull && nfull % 4 == 0) ok = rebase16q<LP, H>(R, total_shift, spread_limit, sbase) && ok;
        run_chunk16q<LP, H, true>(R, penc, dw, lw, sw, p_lo, p_hi, fc, tab_lane, sc1, sc2, lptr, sptr, rem, out, n_st);
    }
    ok = rebase16q<LP, H>(R, total_shift, spread_limit, sbase) && ok;
    const int q_last = (g.sh - 1) - (g.lg - 1) + g.q0;
    const int m_last = q_last % B;
    uint32_t v = 0;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        if (m == m_last) v = R[m] & 0xffffu;
        if (m + H == m_last) v = R[m] >> 16;
    }
    v = __shfl_sync(kFull, v, min(31, sbase + q_last / B));
    if (!ok) return kOverflow16;
    const int x4 = (int)(v & 0x7ffcu) - (int)kTop16 - total_shift;
    return (uint32_t)((int)kBias4 + x4) | 1u;
}

// Trace of a four-pair group's member: `dirw` = the direction words of the group, `seg` = the pair's position in it.
// One lane per pair makes the walk a chain of dependent loads, so every lane walks TWO pairs at once (the loads of both
// are issued before either is used), and each word load sends the sector below it on its way to L2: the walk gets to
// that block of 8 iterations some 8 moves later, near the same diagonal.
struct QWalk {
    const uint32_t *dirw;
    uint32_t *ops;
    int s, q, sft, seg;
    uint32_t n, acc, w;
};
__device__ __forceinline__ void qwalk_init(QWalk &x, const uint32_t *dirw, int seg, int q0, int lg, int sh, uint32_t *ops, bool live)
{
    x.dirw = dirw; x.ops = ops; x.seg = seg;
    x.s = live ? (lg - 1) + (sh - 1) : 0;
    x.q = (sh - 1) - (lg - 1) + q0;
    x.sft = 16; x.n = 0; x.acc = 0; x.w = 0;
}
__device__ __forceinline__ void qwalk_step(QWalk &x)
{
    if (x.s > 0) {
        const uint32_t code = (x.w >> x.sft) & 3u;
        x.acc |= code << (2 * (x.n & 15));
        ++x.n;
        if ((x.n & 15) == 0) { x.ops[(x.n >> 4) - 1] = x.acc; x.acc = 0; }
        if (code == 0u) { x.s -= 2; x.sft += 2; }
        else { x.s -= 1; x.q += (code == 1u) ? 1 : -1; x.sft = 16; }      // 1: consume longer (i--), 2: consume lesser (j--)
    }
}
template <int LP, int H>
__device__ __forceinline__ void trace_two16q(QWalk &a, QWalk &b)
{
    while (a.s > 0 || b.s > 0) {
        const bool la = a.s > 0 && a.sft == 16, lb = b.s > 0 && b.sft == 16;
        // diagonal q = H * h + r: lane h >> 1, register r, half h & 1 -> word H * lane + r of the pair's LP * H
        const int ta = (a.s - (a.q & 1)) >> 1, tb = (b.s - (b.q & 1)) >> 1;
        const uint32_t ha = (uint32_t)a.q / (uint32_t)H, hb = (uint32_t)b.q / (uint32_t)H;
        const uint32_t *pa = a.dirw + ((uint32_t)(ta >> 3) * (32u * H) + (uint32_t)a.seg * (uint32_t)(LP * H) +
                                       (uint32_t)a.q - (uint32_t)H * ((ha + 1u) >> 1));
        const uint32_t *pb = b.dirw + ((uint32_t)(tb >> 3) * (32u * H) + (uint32_t)b.seg * (uint32_t)(LP * H) +
                                       (uint32_t)b.q - (uint32_t)H * ((hb + 1u) >> 1));
        uint32_t wa = 0, wb = 0;
        if (la) wa = __ldcg(pa);
        if (lb) wb = __ldcg(pb);
        if (la && ta >= 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa - (32u * H)));
        if (lb && tb >= 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb - (32u * H)));
        if (la) { a.w = wa >> (16 * (ha & 1u)); a.sft = 2 * (7 - (ta & 7)); }
        if (lb) { b.w = wb >> (16 * (hb & 1u)); b.sft = 2 * (7 - (tb & 7)); }
        qwalk_step(a);
        qwalk_step(b);
        // a second move from the same word where there is one (runs of ALIGN): fewer round trips per alignment
        if (a.sft != 16) qwalk_step(a);
        if (b.sft != 16) qwalk_step(b);
    }
    if (a.n & 15) a.ops[a.n >> 4] = a.acc;
    if (b.n & 15) b.ops[b.n >> 4] = b.acc;
}

// ---------------------------------------------------------------------------------------------
// trace: one lane per pair
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t trace_lane(const uint32_t *dirw, uint32_t wq, int q0, int lg, int sh, uint32_t *ops)
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

// ---------------------------------------------------------------------------------------------
// emit: warp-cooperative, 128 alignment columns per pass
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t emit_pair(const TcmDev &tcm, const uint8_t *pL, const uint8_t *pS,
                                              const uint32_t *ops, uint32_t nal, bool first_long,
                                              uint8_t *og, uint8_t *o1, uint8_t *o2, int lane,
                                              uint8_t *ou = nullptr, uint32_t ucap = 0)
{
    uint32_t base_i = 0, base_j = 0, base_u = 0;
    const uint32_t gap = tcm.gap;
    // The first pass is shortened so that every later one starts on a 128-byte line of the output arrays (their offsets
    // are multiples of 4, and equal for the three): each warp store then covers exactly one whole line — what a posted
    // write over PCIe wants when the arrays are the caller's pinned host buffers, and full sectors in HBM otherwise.
    const uint8_t *oany = og ? og : (o1 ? o1 : o2);
    const int mis = (int)(reinterpret_cast<uintptr_t>(oany) & 124u);
    // the four op codes of this lane's columns in the pass that starts at column c0 (3 = no column), one byte each;
    // fetched one pass ahead: a pass is a chain of dependent loads (ops, then the strings, then the median table)
    auto fetch_codes = [&](int c0) -> uint32_t {
        uint32_t pk = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int cc = c0 + 4 * lane + x;
            uint32_t cd = 3u;
            if (cc >= 0 && cc < (int)nal) {
                const uint32_t r = nal - 1u - (uint32_t)cc;
                cd = (__ldcg(ops + (r >> 4)) >> (2 * (r & 15))) & 3u;
            }
            pk |= cd << (8 * x);
        }
        return pk;
    };
    uint32_t pk_next = fetch_codes(-mis);
    for (int c0 = -mis; c0 < (int)nal; c0 += 128) {
        const int c = c0 + 4 * lane;
        uint32_t code[4];
        uint32_t nl = 0, ns = 0;
        const uint32_t pk = pk_next;
        if (c0 + 128 < (int)nal) pk_next = fetch_codes(c0 + 128);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            code[x] = (pk >> (8 * x)) & 3u;
            nl += (code[x] < 2u);
            ns += (code[x] == 0u || code[x] == 2u);
        }
        const uint32_t packed = nl | (ns << 16);
        uint32_t incl = packed;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t excl = incl - packed;
        uint32_t il = base_i + (excl & 0xffffu), js = base_j + (excl >> 16);
        uint32_t wg = 0, wx = 0, wy = 0, nk = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            uint32_t xx = 0, yy = 0, md = 0;
            if (code[x] != 3u) {
                if (code[x] < 2u) { yy = __ldg(pL + il); ++il; }                    // consume longer
                if (code[x] == 0u || code[x] == 2u) { xx = __ldg(pS + js); ++js; }  // consume lesser
                md = __ldg(tcm.median + (((xx ? xx : gap) << tcm.a) | (yy ? yy : gap)));
                nk += (md != gap);
            }
            wg |= md << (8 * x);
            wx |= xx << (8 * x);
            wy |= yy << (8 * x);
        }
        if (ou) {       // compact the non-gap medians: deleteGaps of the node's context, medians only
            uint32_t inc = nk;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(kFull, inc, d);
                if (lane >= d) inc += o;
            }
            uint32_t pos = base_u + inc - nk;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const uint32_t md = (wg >> (8 * x)) & 0xffu;
                if (code[x] != 3u && md != gap) {
                    if (pos < ucap) ou[pos] = (uint8_t)md;
                    ++pos;
                }
            }
            base_u += __shfl_sync(kFull, inc, 31);
        }
        if (c >= 0 && c < (int)nal) {
            if (og) *reinterpret_cast<uint32_t *>(og + c) = wg;
            if (o1) *reinterpret_cast<uint32_t *>(o1 + c) = first_long ? wx : wy;
            if (o2) *reinterpret_cast<uint32_t *>(o2 + c) = first_long ? wy : wx;
        }
        const uint32_t tot = __shfl_sync(kFull, incl, 31);
        base_i += tot & 0xffffu;
        base_j += tot >> 16;
    }
    return base_u;
}

// Emit for the 16-warp kernel: the same three strings, written without a loop-carried dependency.  emit_pair carries the
// element indices from pass to pass through a warp scan, so every pass is a chain (ops -> scan -> strings -> median ->
// store) that only many resident warps can hide.  Here the op stream is first copied to shared memory together with, per
// 16-move word, the number of longer / lesser elements consumed by all LATER words (= earlier columns); a column's
// element indices are then that count plus a popcount inside its own word, and passes are independent of each other
// (two are kept in flight).  ops_s / suf: scratch in the warp's string windows, which are idle during this phase.
// ou (post-order mode): the medians with the gap elements dropped; only its write positions run from pass to pass (one
// warp scan each), no load address does.  Returns their number.
template <bool UNGAPPED>
__device__ __forceinline__ uint32_t emit_pair_q(const TcmDev &tcm, const uint8_t *med, const uint8_t *pL, const uint8_t *pS,
                                                const uint32_t *ops, uint32_t nal, bool first_long, uint8_t *og, uint8_t *o1,
                                                uint8_t *o2, int lane, uint32_t *ops_s, uint32_t *suf,
                                                uint8_t *ou = nullptr, uint32_t ucap = 0)
{
    const uint32_t gap = tcm.gap;
    const uint32_t nw = (nal + 15u) >> 4, K = (nw + 31u) >> 5;
    uint32_t mine = 0;                                   // (longer | lesser << 16) consumed by this lane's K words
    for (uint32_t x = 0; x < K; ++x) {
        const uint32_t w = (uint32_t)lane * K + x;
        if (w < nw) {
            uint32_t v = __ldcg(ops + w);
            if (w == nw - 1u && (nal & 15u)) v |= 0xffffffffu << (2u * (nal & 15u));      // moves that do not exist: code 3
            ops_s[w] = v;
            mine += (uint32_t)__popc((~v >> 1) & 0x55555555u) | ((uint32_t)__popc(~v & 0x55555555u) << 16);
        }
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_down_sync(kFull, incl, d);
        if (lane + d < 32) incl += o;
    }
    uint32_t run = incl - mine;                          // consumed by the words of all higher lanes
    for (uint32_t x = K; x-- > 0;) {
        const uint32_t w = (uint32_t)lane * K + x;
        if (w < nw) {
            const uint32_t v = ops_s[w];
            suf[w] = run;
            run += (uint32_t)__popc((~v >> 1) & 0x55555555u) | ((uint32_t)__popc(~v & 0x55555555u) << 16);
        }
    }
    __syncwarp();
    const uint8_t *oany = og ? og : (o1 ? o1 : o2);
    const int mis = (int)(reinterpret_cast<uintptr_t>(oany) & 124u);
    uint32_t base_u = 0;
#pragma unroll 2
    for (int c0 = -mis; c0 < (int)nal; c0 += 128) {
        const int c = c0 + 4 * lane;
        const bool mine = c >= 0 && c < (int)nal;
        if (!mine && !UNGAPPED) continue;
        uint32_t xx[4], yy[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            xx[x] = 0; yy[x] = 0;
            const uint32_t r = nal - 1u - (uint32_t)(c + x);              // (wraps for columns past the end: r >= nal)
            if (c + x >= 0 && r < nal) {
                const uint32_t v = ops_s[r >> 4], sh = 2u * (r & 15u), code = (v >> sh) & 3u;
                const uint32_t above = (~v >> sh) >> 2, base = suf[r >> 4];
                const uint32_t il = (base & 0xffffu) + (uint32_t)__popc((above >> 1) & 0x55555555u);
                const uint32_t js = (base >> 16) + (uint32_t)__popc(above & 0x55555555u);
                if (code < 2u) yy[x] = __ldg(pL + il);                    // consume longer
                if (code == 0u || code == 2u) xx[x] = __ldg(pS + js);     // consume lesser
            }
        }
        uint32_t wg = 0, wx = 0, wy = 0, nk = 0, keep = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const uint32_t r = nal - 1u - (uint32_t)(c + x);
            if (c + x >= 0 && r < nal) {
                const uint32_t md = med[((xx[x] ? xx[x] : gap) << tcm.a) | (yy[x] ? yy[x] : gap)];
                wg |= md << (8 * x);
                wx |= xx[x] << (8 * x);
                wy |= yy[x] << (8 * x);
                if (UNGAPPED && md != gap) { ++nk; keep |= 1u << x; }
            }
        }
        if (mine) {
            if (og) *reinterpret_cast<uint32_t *>(og + c) = wg;
            if (o1) *reinterpret_cast<uint32_t *>(o1 + c) = first_long ? wx : wy;
            if (o2) *reinterpret_cast<uint32_t *>(o2 + c) = first_long ? wy : wx;
        }
        if constexpr (UNGAPPED) {       // deleteGaps of the node's context, medians only
            uint32_t inc = nk;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(kFull, inc, d);
                if (lane >= d) inc += o;
            }
            uint32_t pos = base_u + inc - nk;
#pragma unroll
            for (int x = 0; x < 4; ++x)
                if ((keep >> x) & 1u) {
                    if (pos < ucap) ou[pos] = (uint8_t)(wg >> (8 * x));
                    ++pos;
                }
            base_u += __shfl_sync(kFull, inc, 31);
        }
    }
    __syncwarp();
    return base_u;
}

// ---------------------------------------------------------------------------------------------
// per-pair plan: which input is "long" (length, then first differing element), geometry
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool first_is_long(const uint8_t *p1, uint32_t n1, const uint8_t *p2, uint32_t n2, int lane)
{
    if (n1 != n2) return n1 > n2;
    for (uint32_t base = 0; base < n1; base += 32) {
        const uint32_t i = base + lane;
        uint32_t a = 0, b = 0;
        if (i < n1) { a = __ldg(p1 + i); b = __ldg(p2 + i); }
        const unsigned diff = __ballot_sync(kFull, a != b);
        if (diff) {
            const int src = __ffs(diff) - 1;
            const uint32_t aa = __shfl_sync(kFull, a, src), bb = __shfl_sync(kFull, b, src);
            return aa > bb;
        }
    }
    return true;
}

template <int BMAX>
__device__ __forceinline__ uint32_t dispatch_fill(int B, const Geom &g, const FillConst fc, uint32_t tab_lane,
                                                  const char *Lb, const char *Sb, uint32_t *slot, int lane, int sub4gg)
{
    if constexpr (BMAX <= 8) {
        switch (B) {
            case 2:  return fill_pair<2>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg);
            case 4:  return fill_pair<4>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg);
            default: return fill_pair<8>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg);
        }
    } else {
        if (B == 16) return fill_pair<16>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg);
        return fill_pair<32>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg);
    }
}

template <int BMAX>
__device__ __forceinline__ uint32_t dispatch_fill16(int B, const Geom &g, const FillConst fc, uint32_t tab_lane,
                                                    const char *Lb, const char *Sb, uint32_t *slot, int lane, int sub4gg,
                                                    uint32_t spread)
{
    if constexpr (BMAX <= 8) {
        switch (B) {
            case 4:  return fill_pair16<4>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg, spread);
            case 8:  return fill_pair16<8>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg, spread);
            default: return fill_pair16<16>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg, spread);
        }
    } else {
        if (B == 16) return fill_pair16<16>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg, spread);
        return fill_pair16<32>(g, fc, tab_lane, Lb, Sb, slot, lane, sub4gg, spread);
    }
}

// Shared memory: [sub4 table | cgap[dim] | gapc[dim] | per-warp: flags[32], nal[32], off[32], seq window].
extern __shared__ __align__(16) char smem[];

// BMAX = 8 : the narrow kernel (64 registers, up to 32 warps per SM).  Takes pairs straight from the batch;
//            pairs whose band needs 16 or 32 diagonals per lane are queued on big_list.
// BMAX = 32: the wide kernel (128 registers, 16 warps per SM).  Takes its pairs from big_list.
template <int BMAX>
__global__ void __launch_bounds__(BMAX <= 8 ? 1024 : 512, 1) do_linear_kernel(const LinearParams p)
{
    const TcmDev &tcm = p.tcm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int dim = 1 << tcm.a;
    const BatchDev &b = p.b;
    // narrow kernel: straight from the batch, or — after the four-pairs-per-warp kernel — the pairs that one left over
    const bool from_list = BMAX > 8 || p.single_list != nullptr;
    const uint32_t n_work = BMAX <= 8 ? (p.single_list ? *p.single_count : b.n_pairs) : *p.big_count;
    if (n_work == 0) return;

    char *tab = smem;
    uint32_t *s_cgap = reinterpret_cast<uint32_t *>(smem + tcm.sub4_bytes);
    uint32_t *s_gapc = s_cgap + dim;
    char *warp_base = reinterpret_cast<char *>(s_gapc + dim);
    const uint32_t per_warp = 512u + p.seq_cap;
    uint32_t *s_flags = reinterpret_cast<uint32_t *>(warp_base + (size_t)warp * per_warp);
    uint32_t *s_nal = s_flags + 32;
    uint32_t *s_off = s_flags + 64;
    uint32_t *s_pair = s_flags + 96;
    char *seq = reinterpret_cast<char *>(s_flags) + 512;

    for (uint32_t i = threadIdx.x; i < tcm.sub4_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(tab)[i] = __ldg(reinterpret_cast<const uint4 *>(tcm.sub4) + i);
    for (int i = threadIdx.x; i < dim; i += blockDim.x) { s_cgap[i] = tcm.cgap[i]; s_gapc[i] = tcm.gapc[i]; }
    __syncthreads();

    const uint32_t tab_lane = (uint32_t)__cvta_generic_to_shared(tab) + (tcm.repl ? 4u * lane : 0u);
    const FillConst fc = {p.k_four, p.k_minus1, p.l_stride, p.k_64k};
    const int G = p.pairs_per_warp;
    const size_t warp_global = (size_t)blockIdx.x * nwarps + warp;
    uint32_t *arena = p.arena + warp_global * (size_t)p.arena_words;
    unsigned int *counter = BMAX <= 8 ? p.counter : p.counter + 2;
    const bool want_out = b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_len || b.out_ungapped;

    // Work is taken from the counter `grab` pairs at a time (fine-grained: the tail of a batch costs at most one
    // grab per warp) and filled back to back; once `G` pairs are waiting in the arena — or the arena is full, or the
    // batch is exhausted — the warp traces them all at once, one lane per pair, and emits them.
    const uint32_t grab = p.grab ? p.grab : 1u;
    // (Handing the last pairs of a batch out one at a time from a second counter was tried for the many short launches
    // of the host pipeline and the forest: no gain — what a short launch loses is its final trace + emit phase, which all
    // warps reach together and which is latency-bound, not the imbalance of the last grabs.)
    uint32_t gpos = 0, gend = 0;
    bool exhausted = false;
    for (;;) {
        int cnt = 0;
        uint32_t cur = 0;
        // -------------------------------------------------------------- fill, pair after pair
        for (;;) {
            if (cnt == G) break;
            if (gpos == gend) {
                if (exhausted) break;
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(counter, grab);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n_work) { exhausted = true; break; }
                gpos = base;
                gend = min(base + grab, n_work);
            }
            // wide kernel: queue entries carry bit 31 = "the packed fill overflowed its window, use the 32-bit fill"
            const uint32_t entry = !from_list ? gpos : (BMAX <= 8 ? p.single_list[gpos] : p.big_list[gpos]);
            const uint32_t k = entry & 0x7fffffffu;
            const uint32_t n1 = b.len1[k], n2 = b.len2[k];
            const uint8_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
            const bool fl = first_is_long(p1, n1, p2, n2, lane);
            const uint8_t *pL = fl ? p1 : p2, *pS = fl ? p2 : p1;
            const Geom g = make_geom((int)(fl ? n1 : n2) + 1, (int)(fl ? n2 : n1) + 1);
            const int B = choose_b(g.wb);
            // packed 16-bit fill: B = 4 .. 16 in the narrow kernel, B = 16 / 32 in the wide one
            bool packed = tcm.pack_ok && B >= 4 && !(entry >> 31);
            if (BMAX <= 8 && (packed ? (B > 16 || (B == 16 && seq_bytes(seq_layout(g, 16)) > p.seq_cap)) : B > 8)) {
                if (lane == 0) p.big_list[atomicAdd(p.big_count, 1u)] = k;       // -> wide kernel
                ++gpos;
                continue;
            }
            bool ok = B != 0;
            SeqLayout sl = {0, 0, 0, 0};
            uint32_t need = 0;
            if (ok) {
                sl = seq_layout(g, B);
                need = (dir_words(g, B) + ops_words(g) + 3u) & ~3u;
                ok = seq_bytes(sl) <= p.seq_cap && need <= p.arena_words;
            }
            if (!ok) {
                if (lane == 0) {
                    const uint32_t ix = atomicAdd(p.overflow_count, 1u);
                    if (ix < p.overflow_cap) p.overflow_list[ix] = k + b.pair_base;
                }
                ++gpos;
                continue;
            }
            if (cur + need > p.arena_words) break;         // trace + emit what we have, then come back to this pair
            uint32_t *slot = arena + cur;

            // stage both strings; sum the indel costs that are folded out of the cell values
            uint16_t *Sr = reinterpret_cast<uint16_t *>(seq);
            uint8_t *Lr = reinterpret_cast<uint8_t *>(Sr + sl.sizeS);
            uint32_t csum = 0;
            const uint32_t emask = (uint32_t)dim - 1u;
            for (int idx = lane; idx < sl.sizeL; idx += 32) {
                const int i = idx - sl.offL;
                uint32_t e = 0;
                if (i == 0) e = tcm.gap;
                else if (i > 0 && i < g.lg) { e = __ldg(pL + i - 1) & emask; csum += s_cgap[e]; }
                Lr[idx] = (uint8_t)e;
            }
            for (int idx = lane; idx < sl.sizeS; idx += 32) {
                const int j = idx - sl.offS;
                uint32_t e = 0;
                if (j == 0) e = tcm.gap;
                else if (j > 0 && j < g.sh) { e = __ldg(pS + j - 1) & emask; csum += s_gapc[e]; }
                Sr[idx] = (uint16_t)(tcm.repl ? (((e >> 1) << 7) | ((e & 1u) << 1)) : (e << 1));
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(kFull, csum, d);
            __syncwarp();

            const char *Lb = reinterpret_cast<const char *>(Lr + sl.offL);
            const char *Sb = reinterpret_cast<const char *>(Sr + sl.offS);
            uint32_t fw;
            if (packed) {
                fw = dispatch_fill16<BMAX>(B, g, fc, tab_lane, Lb, Sb, slot, lane, tcm.sub4_gapgap, tcm.spread16);
                if (fw == kOverflow16) {                // the sweep's finite spread left the 15-bit window
                    packed = false;
                    if (BMAX <= 8 && B > 8) {           // the 32-bit fill of this width lives in the wide kernel
                        if (lane == 0) p.big_list[atomicAdd(p.big_count, 1u)] = k | 0x80000000u;
                        __syncwarp();
                        ++gpos;
                        continue;
                    }
                    fw = dispatch_fill<BMAX>(B, g, fc, tab_lane, Lb, Sb, slot, lane, tcm.sub4_gapgap);
                }
            } else {
                fw = dispatch_fill<BMAX>(B, g, fc, tab_lane, Lb, Sb, slot, lane, tcm.sub4_gapgap);
            }
            if (lane == 0) {
                s_flags[cnt] = (fl ? 1u : 0u) | (packed ? 4u : 0u);
                s_off[cnt] = cur;
                s_pair[cnt] = k;
                b.cost[k] = (int32_t)((fw >> 2) - (kBias4 >> 2) + csum);
            }
            cur += need;
            if (b.cells) {
                unsigned long long c = band_cells(g, lane, 32);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(kFull, c, d);
                if (lane == 0) b.cells[k] = c;
            }
            ++cnt;
            ++gpos;
            __syncwarp();   // the staging window is reused by the next pair
        }
        __syncwarp();
        if (cnt && want_out) {
            // -------------------------------------------------------------- trace: lane = pair
            if (lane < cnt) {
                const uint32_t k = s_pair[lane];
                const uint32_t n1 = b.len1[k], n2 = b.len2[k];
                const bool fl = s_flags[lane] & 1u;
                const Geom g = make_geom((int)(fl ? n1 : n2) + 1, (int)(fl ? n2 : n1) + 1);
                const int B = choose_b(g.wb);
                uint32_t *slot = arena + s_off[lane];
                const uint32_t n = (s_flags[lane] & 4u) ? trace_lane16(slot, B, g.q0, g.lg, g.sh, slot + dir_words(g, B))
                                                        : trace_lane(slot, 32u * B, g.q0, g.lg, g.sh, slot + dir_words(g, B));
                s_nal[lane] = n;
                if (b.out_len) b.out_len[k] = n;
                if (b.pack_cursor) pack_claim(b, k, n);      // the length is known: claim the pair's packed output range
            }
            __syncwarp();

            // -------------------------------------------------------------- emit: warp per pair
            if (b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_ungapped) {
                for (int pe = 0; pe < cnt; ++pe) {
                    const uint32_t k = s_pair[pe];
                    const uint32_t n1 = b.len1[k], n2 = b.len2[k];
                    const bool fl = s_flags[pe] & 1u;
                    const uint8_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
                    const Geom g = make_geom((int)(fl ? n1 : n2) + 1, (int)(fl ? n2 : n1) + 1);
                    const int B = choose_b(g.wb);
                    const uint32_t *ops = arena + s_off[pe] + dir_words(g, B);
                    const uint64_t oa = b.pack_cursor ? __ldcg(b.out_off_actual + k) : 0;
                    const bool wr = oa != ~0ull;                 // (a packed range that did not fit: see pack_claim)
                    const uint64_t oo = b.pack_cursor ? oa - b.pack_base : b.out_off ? b.out_off[k] : 0;
                    const uint32_t nu = emit_pair(tcm, fl ? p1 : p2, fl ? p2 : p1, ops, s_nal[pe], fl,
                                                  b.out_gapped && wr ? b.out_gapped + oo : nullptr,
                                                  b.out_slot1 && wr ? b.out_slot1 + oo : nullptr,
                                                  b.out_slot2 && wr ? b.out_slot2 + oo : nullptr, lane,
                                                  b.out_ungapped ? b.out_ungapped + b.out_uoff[k] : nullptr, b.ucap);
                    if (b.out_ungapped && lane == 0) {
                        b.out_ulen[k] = nu < b.ucap ? nu : b.ucap;
                        if (nu > b.ucap) atomicAdd(p.overflow_count + 1, 1u);   // node string does not fit its slot
                    }
                }
            }
            __syncwarp();
        }
        if (exhausted && gpos == gend) break;
    }
}

// ---------------------------------------------------------------------------------------------
// Several pairs per warp (fill_multi16<LP, H>: 32 / LP pairs, LP lanes x 2H diagonals each).  Takes groups of consecutive
// pairs — straight from the batch, or from the list an earlier kernel of this family left over; a group all of whose pairs
// have a band this shape is made for, fit the staging windows and need about the same number of iterations is swept at
// once, every other pair goes to the output list for the next kernel (another shape, then the narrow kernel).
// Shared memory: [sub4 table | cgap | gapc | per warp: group / pair records (640 bytes) + 32 / LP string windows].
// ---------------------------------------------------------------------------------------------
template <int LP>
__device__ __forceinline__ bool first_is_long_seg(const uint8_t *p1, uint32_t n1, const uint8_t *p2, uint32_t n2, int lane,
                                                  int ln, int sb)
{
    bool decided = n1 != n2 || n1 == 0, res = n1 >= n2;
    for (uint32_t base = 0; __any_sync(kFull, !decided); base += LP) {
        const uint32_t i = base + ln;
        uint32_t a = 0, b = 0;
        if (!decided && i < n1) { a = __ldg(p1 + i); b = __ldg(p2 + i); }
        const unsigned diff = (__ballot_sync(kFull, a != b) >> sb) & ((1u << LP) - 1u);
        const int src = diff ? sb + __ffs(diff) - 1 : lane;
        const uint32_t aa = __shfl_sync(kFull, a, src), bb = __shfl_sync(kFull, b, src);
        if (!decided) {
            if (diff) { res = aa > bb; decided = true; }
            else if (base + LP >= n1) { res = true; decided = true; }
        }
    }
    return res;
}

// PROFILE: per-phase cycle counts into p.phase_cycles (an instantiation of its own: the counters cost registers)
template <int LP, int H, bool PROFILE>
__global__ void __launch_bounds__(512, 1) do_linear_multi_kernel(const LinearParams p)
{
    constexpr int NP = 32 / LP, B = 2 * H;
    // bands this shape takes: more diagonals than the next smaller shape (or the narrow kernel's B = 2) holds
    constexpr int kLow = LP == 8 ? 6 * 28 : (LP == 6 ? 5 * 28 : (LP == 5 ? 128 : 64));
    const TcmDev &tcm = p.tcm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // lanes behind the last pair (LP = 5: lanes 30, 31) shadow the first lanes of the last pair: same loads, same window
    // writes, results unused
    const bool real = lane < NP * LP;
    const int seg = real ? lane / LP : NP - 1, ln = lane % LP;
    const bool lead = real && ln == 0;
    const int dim = 1 << tcm.a;
    const BatchDev &b = p.b;
    const uint32_t n_work = p.multi_in_list ? *p.multi_in_count : b.n_pairs;
    if (n_work == 0) return;
    const uint32_t *in_list = p.multi_in_list;

    char *tab = smem;
    uint32_t *s_cgap = reinterpret_cast<uint32_t *>(smem + tcm.sub4_bytes);
    uint32_t *s_gapc = s_cgap + dim;
    char *warp_base = reinterpret_cast<char *>(s_gapc + dim);
    const uint32_t per_warp = 640u + NP * p.seq_cap;
    // per group (16): arena offset of its direction words, of its op streams, words per op stream; per pair (64): index
    // in the batch (bit 31: the first input is the longer one, bit 30: left to the narrow kernel) and alignment length
    uint32_t *g_off = reinterpret_cast<uint32_t *>(warp_base + (size_t)warp * per_warp);
    uint32_t *g_ops = g_off + 16, *g_opw = g_off + 32, *s_pair = g_off + 48;
    uint16_t *s_nal = reinterpret_cast<uint16_t *>(g_off + 112);
    // (each pair's window starts one bank further than the previous one's: with equal geometries the segments' window
    // loads would otherwise hit the same banks)
    uint8_t *win = reinterpret_cast<uint8_t *>(g_off) + 640 + (size_t)seg * p.seq_cap + 4 * seg;

    for (uint32_t i = threadIdx.x; i < tcm.sub4_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(tab)[i] = __ldg(reinterpret_cast<const uint4 *>(tcm.sub4) + i);
    for (int i = threadIdx.x; i < dim; i += blockDim.x) { s_cgap[i] = tcm.cgap[i]; s_gapc[i] = tcm.gapc[i]; }
    __syncthreads();

    const bool repl = tcm.repl != 0;
    const uint32_t tab_lane = (uint32_t)__cvta_generic_to_shared(tab) + (repl ? 4u * lane : 0u);
    const FillConst fc = {p.k_four, p.k_minus1, p.l_stride, p.k_64k};
    int G = (2 * p.pairs_per_warp) / NP * NP;          // two walks per lane in the trace
    if (G < NP) G = NP;
    // The warps of a CTA start together and a uniform batch keeps them in step: all fill (issue-bound) and then all trace
    // (latency-bound) at the same time.  Experiment (quad_flags bit 0): a first batch of a different size per warp spreads
    // the phases for good — measured SLOWER (67.5 -> 77.9 ms on config 2), so off by default.
    int Gcur = (p.quad_flags & 1u) ? NP * (1 + (warp * (G / NP - 1)) / (nwarps > 1 ? nwarps - 1 : 1)) : G;
    long long ph_t = PROFILE ? clock64() : 0;
    unsigned long long ph[6] = {0, 0, 0, 0, 0, 0};
    auto phase = [&](int i) { if constexpr (PROFILE) { const long long t = clock64(); ph[i] += (unsigned long long)(t - ph_t); ph_t = t; } };
    const size_t warp_global = (size_t)blockIdx.x * nwarps + warp;
    uint32_t *arena = p.arena + warp_global * (size_t)p.quad_arena_words;
    unsigned int *counter = p.multi_counter;
    const bool want_out = b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_len || b.out_ungapped;
    const uint32_t emask = (uint32_t)dim - 1u;

    constexpr uint32_t kNone = 0xffffffffu;
    uint32_t pend = kNone, nxt = kNone;
    bool exhausted = false;
    for (;;) {
        int cnt = 0;
        uint32_t cur = 0;
        // -------------------------------------------------------------- fill, group after group
        for (;;) {
            if (cnt + NP > Gcur) break;
            if (pend == kNone) {
                if (nxt != kNone) { pend = nxt; nxt = kNone; }
                else {
                    if (exhausted) break;
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(counter, (unsigned int)NP);
                    base = __shfl_sync(kFull, base, 0);
                    if (base >= n_work) { exhausted = true; break; }
                    pend = base;
                }
            }
            if (nxt == kNone && !exhausted) {
                // take the group after this one now and start its strings on their way from HBM to L2
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(counter, (unsigned int)NP);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n_work) exhausted = true;
                else {
                    nxt = base;
                    const uint32_t kx = base + (uint32_t)seg;
                    if (kx < n_work) {
                        const uint32_t kn = in_list ? (in_list[kx] & 0x7fffffffu) : kx;
                        const uint8_t *q1 = b.elems + b.off1[kn], *q2 = b.elems + b.off2[kn];
                        const uint32_t m1 = b.len1[kn], m2 = b.len2[kn];
                        for (uint32_t o = 128u * ln; m1 && o < m1 + 127u; o += 128u * LP)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + min(o, m1 - 1u)));
                        for (uint32_t o = 128u * ln; m2 && o < m2 + 127u; o += 128u * LP)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + min(o, m2 - 1u)));
                    }
                }
            }
            const bool have = pend + (uint32_t)seg < n_work;
            const uint32_t kx = have ? pend + (uint32_t)seg : pend;
            const uint32_t entry = in_list ? in_list[kx] : kx;          // (bit 31: already marked for the 32-bit fill)
            const uint32_t k = entry & 0x7fffffffu;
            const uint32_t n1 = b.len1[k], n2 = b.len2[k];
            const uint8_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
            const bool fl = first_is_long_seg<LP>(p1, n1, p2, n2, lane, ln, seg * LP);
            const uint8_t *pL = fl ? p1 : p2, *pS = fl ? p2 : p1;
            const Geom g = make_geom((int)(fl ? n1 : n2) + 1, (int)(fl ? n2 : n1) + 1);
            int tmax = g.T;
            uint32_t opw = ops_words(g);
            tmax = __reduce_max_sync(kFull, tmax);
            opw = __reduce_max_sync(kFull, opw);
            const SeqLayout sl = seq_layout_multi<LP, H>(g, tmax);
            const uint32_t dirw = multi_dir_words<H>(tmax);
            const uint32_t need = (dirw + NP * opw + 3u) & ~3u;
            // the group's widest band decides the shape (a narrower member rides along at this shape's width)
            const int wmax = __reduce_max_sync(kFull, have ? g.wb + 1 : 0);
            const bool elig = have && !(entry >> 31) && wmax <= LP * B && wmax > kLow && g.sh > 8 &&
                              (uint32_t)(sl.sizeL + sl.sizeS) + 4u * NP <= p.seq_cap && (tmax - g.T) * 8 <= tmax &&
                              need <= p.quad_arena_words && g.lg + g.sh < 65536 && k < (1u << 30);
            if (!__all_sync(kFull, elig)) {
                // handed on as a block: the next kernel forms its groups from consecutive list entries, and neighbours in
                // the batch are what is likely to be of one size
                const uint32_t n_have = min((uint32_t)NP, n_work - pend);
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(p.single_count, n_have);
                at = __shfl_sync(kFull, at, 0);
                if (lead && have) p.single_list[at + (uint32_t)seg] = entry;
                pend = kNone;
                continue;
            }
            if (cur + need > p.quad_arena_words) break;    // trace + emit what we have, then come back to this group
            uint32_t *slot = arena + cur;

            // stage the segment's pair; sum the indel costs that are folded out of the cell values
            // (four bytes per lane and step, four steps in flight; the loads are unconditional — clamped addresses — because
            // ptxas keeps a predicated load next to its use, and these come from L2 / HBM with only 16 warps to hide them)
            uint8_t *Lr = win, *Sr = win + sl.sizeL;
            uint32_t csum = 0;
            phase(4);
            {
                const int nL = g.lg - 1, nS = g.sh - 1;          // >= 8 (eligibility)
#pragma unroll 4
                for (int w4 = ln; w4 < sl.sizeL / 4; w4 += LP) {
                    uint32_t e[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int i = 4 * w4 + x - sl.offL - 1;
                        e[x] = __ldg(pL + min(max(i, 0), nL - 1));
                    }
                    uint32_t word = 0;
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int i = 4 * w4 + x - sl.offL - 1;
                        uint32_t v = e[x] & emask;
                        if (i >= 0 && i < nL) csum += s_cgap[v];
                        else v = (i == -1) ? tcm.gap : 0u;
                        word |= v << (8 * x);
                    }
                    reinterpret_cast<uint32_t *>(Lr)[w4] = word;
                }
#pragma unroll 4
                for (int w4 = ln; w4 < sl.sizeS / 4; w4 += LP) {
                    uint32_t e[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int j = 4 * w4 + x - sl.offS - 1;
                        e[x] = __ldg(pS + min(max(j, 0), nS - 1));
                    }
                    uint32_t word = 0;
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int j = 4 * w4 + x - sl.offS - 1;
                        uint32_t v = e[x] & emask;
                        if (j >= 0 && j < nS) csum += s_gapc[v];
                        else v = (j == -1) ? tcm.gap : 0u;
                        word |= v << (8 * x);
                    }
                    reinterpret_cast<uint32_t *>(Sr)[w4] = word;
                }
            }
            csum = seg_reduce<LP>(csum, seg * LP, [](uint32_t x, uint32_t y) { return x + y; });
            __syncwarp();
            phase(0);

            const uint32_t fw = fill_multi16<LP, H>(g, tmax, fc, tab_lane, repl, reinterpret_cast<const char *>(Lr + sl.offL),
                                            reinterpret_cast<const char *>(Sr + sl.offS), slot, lane, tcm.sub4_gapgap,
                                            tcm.spread16);
            phase(1);
            const bool over = fw == kOverflow16;           // the sweep's finite spread left the 15-bit window:
            if (lead) {                                    //   the narrow kernel redoes the pair with the 32-bit fill
                if (over) p.single_list[atomicAdd(p.single_count, 1u)] = k | 0x80000000u;
                else b.cost[k] = (int32_t)((fw >> 2) - (kBias4 >> 2) + csum);
                s_pair[cnt + seg] = k | (fl ? 0x80000000u : 0u) | (over ? 0x40000000u : 0u);
            }
            if (lane == 0) { g_off[cnt / NP] = cur; g_ops[cnt / NP] = cur + dirw; g_opw[cnt / NP] = opw; }
            if (b.cells && !over) {
                const unsigned long long c = band_cells(g, ln, LP);
                const uint32_t c_lo = seg_reduce<LP>((uint32_t)(c & 0xffffffu), seg * LP, [](uint32_t x, uint32_t y) { return x + y; });
                const uint32_t c_hi = seg_reduce<LP>((uint32_t)(c >> 24), seg * LP, [](uint32_t x, uint32_t y) { return x + y; });
                if (lead) b.cells[k] = ((unsigned long long)c_hi << 24) + c_lo;
            }
            cur += need;
            cnt += NP;
            pend = kNone;
            __syncwarp();   // the staging windows are reused by the next group
            phase(5);
        }
        __syncwarp();
        if (cnt && want_out) {
            // -------------------------------------------------------------- trace: lane = pairs `lane` and `lane + 32`
            phase(5);
            {
                QWalk wk[2];
                uint32_t kk[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ix = lane + 32 * h;
                    const bool live = ix < cnt && !(s_pair[ix < cnt ? ix : 0] & 0x40000000u);
                    const uint32_t e = s_pair[live ? ix : 0];
                    const uint32_t k = e & 0x3fffffffu;
                    kk[h] = live ? k : 0xffffffffu;
                    const uint32_t n1 = b.len1[k], n2 = b.len2[k];
                    const bool fl = e >> 31;
                    const Geom g = make_geom((int)(fl ? n1 : n2) + 1, (int)(fl ? n2 : n1) + 1);
                    const int gi = (live ? ix : 0) / NP;
                    qwalk_init(wk[h], arena + g_off[gi], ix % NP, g.q0, g.lg, g.sh, arena + g_ops[gi] + (uint32_t)(ix % NP) * g_opw[gi], live);
                }
                trace_two16q<LP, H>(wk[0], wk[1]);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (kk[h] != 0xffffffffu) {
                        s_nal[lane + 32 * h] = (uint16_t)wk[h].n;
                        if (b.out_len) b.out_len[kk[h]] = wk[h].n;
                        if (b.pack_cursor) pack_claim(b, kk[h], wk[h].n);
                    }
                }
            }
            __syncwarp();
            phase(2);

            // -------------------------------------------------------------- emit: warp per pair
            if (b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_ungapped) {
                // scratch in the (idle) string windows: op stream of the pair, per-word element counts, the median table
                uint32_t *ops_s = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(g_off) + 640);
                const uint32_t nws = p.seq_cap / 16u + 4u;
                uint32_t *suf = ops_s + nws;
                const uint8_t *med = tcm.median;
                if ((uint32_t)(dim * dim) + 8u * nws <= NP * p.seq_cap) {
                    uint32_t *ms = suf + nws;
                    for (int i = lane; i < dim * dim / 4; i += 32) ms[i] = __ldg(reinterpret_cast<const uint32_t *>(tcm.median) + i);
                    med = reinterpret_cast<const uint8_t *>(ms);
                    __syncwarp();
                }
                for (int pe = 0; pe < cnt; ++pe) {
                    const uint32_t e = s_pair[pe];
                    if (e & 0x40000000u) continue;
                    const uint32_t k = e & 0x3fffffffu;
                    const bool fl = e >> 31;
                    const uint8_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
                    const uint64_t oa = b.pack_cursor ? __ldcg(b.out_off_actual + k) : 0;
                    const bool wr = oa != ~0ull;
                    const uint64_t oo = b.pack_cursor ? oa - b.pack_base : b.out_off ? b.out_off[k] : 0;
                    const uint32_t *ops = arena + g_ops[pe / NP] + (uint32_t)(pe % NP) * g_opw[pe / NP];
                    if (pe + 1 < cnt) {          // the next pair's strings: staged long ago, on their way back to L2 now
                        const uint32_t kn = s_pair[pe + 1] & 0x3fffffffu;
                        const uint8_t *q = b.elems + ((lane & 16) ? b.off2[kn] : b.off1[kn]);
                        const uint32_t m = (lane & 16) ? b.len2[kn] : b.len1[kn];
                        const uint32_t o = 128u * (uint32_t)(lane & 15);
                        if (m && o < m + 127u) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + min(o, m - 1u)));
                    }
                    uint32_t nu = 0;
                    if (b.out_ungapped)
                        nu = emit_pair_q<true>(tcm, med, fl ? p1 : p2, fl ? p2 : p1, ops, s_nal[pe], fl,
                                               b.out_gapped && wr ? b.out_gapped + oo : nullptr,
                                               b.out_slot1 && wr ? b.out_slot1 + oo : nullptr,
                                               b.out_slot2 && wr ? b.out_slot2 + oo : nullptr, lane, ops_s, suf,
                                               b.out_ungapped + b.out_uoff[k], b.ucap);
                    else if (wr)
                        emit_pair_q<false>(tcm, med, fl ? p1 : p2, fl ? p2 : p1, ops, s_nal[pe], fl,
                                           b.out_gapped ? b.out_gapped + oo : nullptr, b.out_slot1 ? b.out_slot1 + oo : nullptr,
                                           b.out_slot2 ? b.out_slot2 + oo : nullptr, lane, ops_s, suf);
                    if (b.out_ungapped && lane == 0) {
                        b.out_ulen[k] = nu < b.ucap ? nu : b.ucap;
                        if (nu > b.ucap) atomicAdd(p.overflow_count + 1, 1u);   // node string does not fit its slot
                    }
                }
            }
            __syncwarp();
            phase(3);
        }
        Gcur = G;
        if (exhausted && pend == kNone && nxt == kNone) break;
    }
    if (PROFILE && lane == 0) {
        phase(4);
        for (int i = 0; i < 6; ++i) atomicAdd(p.phase_cycles + i, ph[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// General-geometry kernel: one CTA per pair, any band width / string length.  Same recurrence, same
// direction-word layout (so trace_lane / emit_pair are shared), but the diagonal values live in global
// memory and each anti-diagonal is swept by the whole CTA with a barrier in between.  Used only for the
// pairs the warp kernels cannot hold (band wider than 1023 diagonals, strings beyond the shared-memory
// window): correct for every input, not tuned.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t general_wq(const Geom &g) { return (uint32_t)(g.wb + 1 + 3) & ~3u; }
__host__ __device__ inline uint64_t general_words(const Geom &g)
{
    const uint64_t wq = general_wq(g);
    return (wq + 8) + (uint64_t)((g.T + 15) / 16) * wq + ops_words(g) + 8;
}

// The diagonal values in global memory, direction words by read-modify-write: for what do_general_kernel cannot hold in
// shared memory / registers (bands beyond 8192 diagonals).
__device__ __noinline__ void general_pair_global(const LinearParams &p, const uint32_t *list, const uint64_t *scratch_off,
                                                 uint32_t *scratch)
{
    const TcmDev &tcm = p.tcm;
    const BatchDev &b = p.b;
    const uint32_t k = list[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_fl;
    __shared__ uint32_t s_red[32];
    __shared__ uint32_t s_n;
    const uint32_t n1 = b.len1[k], n2 = b.len2[k];
    const uint8_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
    if (warp == 0) {
        const bool f = first_is_long(p1, n1, p2, n2, lane);
        if (lane == 0) s_fl = f;
    }
    __syncthreads();
    const bool fl = s_fl != 0;
    const uint8_t *pL = fl ? p1 : p2, *pS = fl ? p2 : p1;
    const Geom g = make_geom((int)(fl ? n1 : n2) + 1, (int)(fl ? n2 : n1) + 1);
    const uint32_t wq = general_wq(g);
    uint32_t *w = scratch + scratch_off[blockIdx.x];          // w[1 + q], fences at 0 and wq + 1
    uint32_t *dirw = w + wq + 8;
    const uint32_t ndir = (uint32_t)((g.T + 15) / 16) * wq;
    uint32_t *ops = dirw + ndir;
    for (uint32_t i = threadIdx.x; i < wq + 8; i += blockDim.x) w[i] = kInfW;
    for (uint32_t i = threadIdx.x; i < ndir; i += blockDim.x) dirw[i] = 0;
    uint32_t csum = 0;
    const uint32_t emask = (1u << tcm.a) - 1u;
    for (int i = 1 + threadIdx.x; i < g.lg; i += blockDim.x) csum += tcm.cgap[pL[i - 1] & emask];
    for (int j = 1 + threadIdx.x; j < g.sh; j += blockDim.x) csum += tcm.gapc[pS[j - 1] & emask];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(kFull, csum, d);
    if (lane == 0) s_red[warp] = csum;
    __syncthreads();
    csum = 0;
    for (int x = 0; x < (int)(blockDim.x >> 5); ++x) csum += s_red[x];
    if (threadIdx.x == 0) w[1 + g.q0] = kBias4 - (uint32_t)tcm.sub4_gapgap;
    __syncthreads();

    const int smax = (g.lg - 1) + (g.sh - 1);
    for (int s = 0; s <= smax; ++s) {
        const int i_lo = s - (g.sh - 1) > 0 ? s - (g.sh - 1) : 0;
        const int i_hi = s < g.lg - 1 ? s : g.lg - 1;
        for (int i = i_lo + (int)threadIdx.x; i <= i_hi; i += blockDim.x) {
            const int j = s - i;
            const int q = j - i + g.q0;
            if (q < g.pad || q >= g.wb) continue;
            const uint32_t l = i == 0 ? tcm.gap : (pL[i - 1] & emask);
            const uint32_t sj = j == 0 ? tcm.gap : (pS[j - 1] & emask);
            const uint32_t sub = (uint32_t)(int)tcm.sub4_nat[(l << tcm.a) | sj];
            const uint32_t t1 = __viaddmin_u32(w[1 + q], sub, w[2 + q]);
            const uint32_t mn = __viaddmin_u32(w[q], 1u, t1);
            const int t = (s - (q & 1)) >> 1;
            dirw[(uint32_t)(t >> 4) * wq + (uint32_t)q] |= (mn & 3u) << (2 * (15 - (t & 15)));
            w[1 + q] = (mn & ~3u) | 1u;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int q_last = (g.sh - 1) - (g.lg - 1) + g.q0;
        b.cost[k] = (int32_t)((w[1 + q_last] >> 2) - (kBias4 >> 2) + csum);
        if (b.cells) b.cells[k] = band_cells(g, 0, 1);
        const bool want_out = b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_len || b.out_ungapped;
        uint32_t n = 0;
        if (want_out) {
            n = trace_lane(dirw, wq, g.q0, g.lg, g.sh, ops);
            if (b.out_len) b.out_len[k] = n;
            if (b.pack_cursor) pack_claim(b, k, n);
        }
        s_n = n;
        __threadfence_block();
    }
    __syncthreads();
    if (warp == 0 && (b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_ungapped)) {
        const uint64_t oa = b.pack_cursor ? __ldcg(b.out_off_actual + k) : 0;
        const bool wr = oa != ~0ull;
        const uint64_t oo = b.pack_cursor ? oa - b.pack_base : b.out_off ? b.out_off[k] : 0;
        const uint32_t nu = emit_pair(tcm, pL, pS, ops, s_n, fl, b.out_gapped && wr ? b.out_gapped + oo : nullptr,
                                      b.out_slot1 && wr ? b.out_slot1 + oo : nullptr,
                                      b.out_slot2 && wr ? b.out_slot2 + oo : nullptr, lane,
                                      b.out_ungapped ? b.out_ungapped + b.out_uoff[k] : nullptr, b.ucap);
        if (b.out_ungapped && lane == 0) {
            b.out_ulen[k] = nu < b.ucap ? nu : b.ucap;
            if (nu > b.ucap) atomicAdd(p.overflow_count + 1, 1u);
        }
    }
}

// Tuned path (round 2): the reference's own CPU benchmark (config 1) is 190 pairs of 8 .. 4608 elements, half of them in
// full-matrix mode with 1000 .. 6900 diagonals — one CTA each, so a launch lasts as long as its longest pair: 6900
// anti-diagonal steps.  A step used to be a round of global-memory loads + a read-modify-write of the direction word
// (2 µs); here the diagonal values, both strings and the cost table live in shared memory, every thread owns up to 8 fixed
// diagonals (q = tid, tid + 1024, ...) whose direction words it builds in registers and stores once per 16 cells, and a
// step is a few dozen instructions and one barrier.
constexpr int kGenThreads = 1024, kGenOwn = 8;

__global__ void __launch_bounds__(kGenThreads) do_general_kernel(const LinearParams p, const uint32_t *list,
                                                                 const uint64_t *scratch_off, uint32_t *scratch, uint32_t smem_bytes)
{
    const TcmDev &tcm = p.tcm;
    const BatchDev &b = p.b;
    const uint32_t k = list[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    __shared__ int s_fl;
    __shared__ uint32_t s_red[32];
    __shared__ uint32_t s_n;
    const uint32_t n1 = b.len1[k], n2 = b.len2[k];
    const uint8_t *p1 = b.elems + b.off1[k], *p2 = b.elems + b.off2[k];
    if (warp == 0) {
        const bool f = first_is_long(p1, n1, p2, n2, lane);
        if (lane == 0) s_fl = f;
    }
    __syncthreads();
    const bool fl = s_fl != 0;
    const uint8_t *pL = fl ? p1 : p2, *pS = fl ? p2 : p1;
    const Geom g = make_geom((int)(fl ? n1 : n2) + 1, (int)(fl ? n2 : n1) + 1);
    const uint32_t wq = general_wq(g);
    const int dim = 1 << tcm.a;
    const uint32_t tab_bytes = (uint32_t)(dim * dim * 2);
    const bool tab_in = tcm.a <= 7;
    const uint32_t need = (wq + 8u) * 4u + (tab_in ? tab_bytes : 0u) + (((uint32_t)g.lg + 3u) & ~3u) + (((uint32_t)g.sh + 3u) & ~3u);
    if (need > smem_bytes || (uint32_t)g.wb > (uint32_t)(kGenThreads * kGenOwn)) {      // (uniform over the CTA)
        general_pair_global(p, list, scratch_off, scratch);
        return;
    }
    uint32_t *w = reinterpret_cast<uint32_t *>(smem);                    // w[1 + q], fences at 0 and wq + 1
    int16_t *tab = reinterpret_cast<int16_t *>(w + wq + 8);
    uint8_t *Ls = reinterpret_cast<uint8_t *>(tab) + (tab_in ? tab_bytes : 0u);
    uint8_t *Ss = Ls + (((uint32_t)g.lg + 3u) & ~3u);
    uint32_t *gbase = scratch + scratch_off[blockIdx.x];
    uint32_t *dirw = gbase + wq + 8;                                     // (same scratch layout as the global path)
    const uint32_t ndir = (uint32_t)((g.T + 15) / 16) * wq;
    uint32_t *ops = dirw + ndir;
    const uint32_t emask = (uint32_t)dim - 1u;
    for (uint32_t i = tid; i < wq + 8; i += kGenThreads) w[i] = kInfW;
    if (tab_in) for (int i = tid; i < dim * dim; i += kGenThreads) tab[i] = tcm.sub4_nat[i];
    uint32_t csum = 0;
    for (int i = tid; i < g.lg; i += kGenThreads) {
        uint32_t e = tcm.gap;
        if (i > 0) { e = pL[i - 1] & emask; csum += tcm.cgap[e]; }
        Ls[i] = (uint8_t)e;
    }
    for (int j = tid; j < g.sh; j += kGenThreads) {
        uint32_t e = tcm.gap;
        if (j > 0) { e = pS[j - 1] & emask; csum += tcm.gapc[e]; }
        Ss[j] = (uint8_t)e;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) csum += __shfl_xor_sync(kFull, csum, d);
    if (lane == 0) s_red[warp] = csum;
    __syncthreads();
    csum = 0;
    for (int x = 0; x < kGenThreads / 32; ++x) csum += s_red[x];
    if (tid == 0) w[1 + g.q0] = kBias4 - (uint32_t)tcm.sub4_gapgap;
    __syncthreads();

    // Thread tid owns the diagonals q = 2 * (tid + o * 1024) + c, o < 4, of BOTH parities c: a step only touches the
    // diagonals of its own parity (q0 is even), so every thread has work in every step and no slot is looked at in vain.
    uint32_t acc[kGenOwn];
#pragma unroll
    for (int o = 0; o < kGenOwn; ++o) acc[o] = 0;
    const int16_t *tb = tab_in ? tab : tcm.sub4_nat;
    const int smax = (g.lg - 1) + (g.sh - 1);
    for (int s = 0; s <= smax; ++s) {
        const int c = s & 1;
        // diagonals with a cell on anti-diagonal s: i = (s - d) / 2 in [0, lg), j = (s + d) / 2 in [0, sh), inside the band
        const int q_lo = max(g.pad, g.q0 + max(s - 2 * (g.lg - 1), -s)), q_hi = min(g.wb - 1, g.q0 + min(s, 2 * (g.sh - 1) - s));
#pragma unroll
        for (int oo = 0; oo < kGenOwn / 2; ++oo) {
            const int o = 2 * oo + c;            // (accumulator slot; resolved by predication below)
            const int q = 2 * (tid + oo * kGenThreads) + c;
            if (q < q_lo || q > q_hi) continue;
            const int d = q - g.q0;
            const int i = (s - d) >> 1, j = i + d;
            const uint32_t sub = (uint32_t)(int)tb[((uint32_t)Ls[i] << tcm.a) | Ss[j]];
            const uint32_t t1 = __viaddmin_u32(w[1 + q], sub, w[2 + q]);
            const uint32_t mn = __viaddmin_u32(w[q], 1u, t1);
            const int t = (s - (q & 1)) >> 1;
            uint32_t a = (c ? acc[2 * oo + 1] : acc[2 * oo]) | ((mn & 3u) << (2 * (15 - (t & 15))));
            w[1 + q] = (mn & ~3u) | 1u;
            // the word is complete, or this was the diagonal's last cell
            if ((t & 15) == 15 || i == g.lg - 1 || j == g.sh - 1) {
                dirw[(uint32_t)(t >> 4) * wq + (uint32_t)q] = a;
                if ((t & 15) == 15) a = 0;
            }
            if (c) acc[2 * oo + 1] = a; else acc[2 * oo] = a;
            (void)o;
        }
        __syncthreads();
    }
    if (tid == 0) {
        const int q_last = (g.sh - 1) - (g.lg - 1) + g.q0;
        b.cost[k] = (int32_t)((w[1 + q_last] >> 2) - (kBias4 >> 2) + csum);
        if (b.cells) b.cells[k] = band_cells(g, 0, 1);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const bool want_out = b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_len || b.out_ungapped;
        uint32_t n = 0;
        if (want_out) {
            n = trace_lane(dirw, wq, g.q0, g.lg, g.sh, ops);
            if (b.out_len) b.out_len[k] = n;
            if (b.pack_cursor) pack_claim(b, k, n);
        }
        s_n = n;
        __threadfence_block();
    }
    __syncthreads();
    if (warp == 0 && (b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_ungapped)) {
        const uint64_t oa = b.pack_cursor ? __ldcg(b.out_off_actual + k) : 0;
        const bool wr = oa != ~0ull;
        const uint64_t oo = b.pack_cursor ? oa - b.pack_base : b.out_off ? b.out_off[k] : 0;
        const uint32_t nu = emit_pair(tcm, pL, pS, ops, s_n, fl, b.out_gapped && wr ? b.out_gapped + oo : nullptr,
                                      b.out_slot1 && wr ? b.out_slot1 + oo : nullptr,
                                      b.out_slot2 && wr ? b.out_slot2 + oo : nullptr, lane,
                                      b.out_ungapped ? b.out_ungapped + b.out_uoff[k] : nullptr, b.ucap);
        if (b.out_ungapped && lane == 0) {
            b.out_ulen[k] = nu < b.ucap ? nu : b.ucap;
            if (nu > b.ucap) atomicAdd(p.overflow_count + 1, 1u);
        }
    }
}

uint64_t general_words_for(uint32_t n1, uint32_t n2)
{
    const uint32_t lg = (n1 > n2 ? n1 : n2) + 1, sh = (n1 > n2 ? n2 : n1) + 1;
    return (general_words(make_geom((int)lg, (int)sh)) + 3) & ~(uint64_t)3;
}

cudaError_t launch_general(const LinearParams &p, uint32_t n_list, const uint32_t *list, const uint64_t *scratch_off,
                           uint32_t *scratch, cudaStream_t stream)
{
    constexpr uint32_t kSmem = 200 * 1024;
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(do_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
        once.mark(dev);
    }
    do_general_kernel<<<n_list, kGenThreads, kSmem, stream>>>(p, list, scratch_off, scratch, kSmem);
    return cudaGetLastError();
}

// Host-side launchers (called from do_api.cu).
// shape 0: 8 lanes x 28 diagonals (4 pairs per warp), 1: 4 lanes x 32 (8 pairs), 2: 5 lanes x 28 (6 pairs, 2 lanes idle),
// 3: 6 lanes x 28 (5 pairs, 2 lanes idle)
cudaError_t launch_linear_multi(const LinearParams &p, int shape, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        const void *fns[] = {(const void *)do_linear_multi_kernel<8, 14, false>, (const void *)do_linear_multi_kernel<8, 14, true>,
                             (const void *)do_linear_multi_kernel<4, 16, false>, (const void *)do_linear_multi_kernel<4, 16, true>,
                             (const void *)do_linear_multi_kernel<5, 14, false>, (const void *)do_linear_multi_kernel<5, 14, true>,
                             (const void *)do_linear_multi_kernel<6, 14, false>, (const void *)do_linear_multi_kernel<6, 14, true>};
        for (const void *f : fns) {
            cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return e;
        }
        once.mark(dev);
    }
    if (shape == 0) {
        if (p.phase_cycles) do_linear_multi_kernel<8, 14, true><<<grid, block, smem_bytes, stream>>>(p);
        else                do_linear_multi_kernel<8, 14, false><<<grid, block, smem_bytes, stream>>>(p);
    } else if (shape == 1) {
        if (p.phase_cycles) do_linear_multi_kernel<4, 16, true><<<grid, block, smem_bytes, stream>>>(p);
        else                do_linear_multi_kernel<4, 16, false><<<grid, block, smem_bytes, stream>>>(p);
    } else if (shape == 2) {
        if (p.phase_cycles) do_linear_multi_kernel<5, 14, true><<<grid, block, smem_bytes, stream>>>(p);
        else                do_linear_multi_kernel<5, 14, false><<<grid, block, smem_bytes, stream>>>(p);
    } else {
        if (p.phase_cycles) do_linear_multi_kernel<6, 14, true><<<grid, block, smem_bytes, stream>>>(p);
        else                do_linear_multi_kernel<6, 14, false><<<grid, block, smem_bytes, stream>>>(p);
    }
    return cudaGetLastError();
}

// staging bytes one pair of a multi-pair group may need for strings of up to max_len elements (shape as above)
uint32_t multi_seq_cap_for(int max_len, int shape)
{
    const int lanes_h = shape == 0 ? 8 * 14 : (shape == 1 ? 4 * 16 : (shape == 2 ? 5 * 14 : 6 * 14));
    const int tpad = (max_len + 1 + 15) & ~15;
    const int q0h = (shape == 0 ? 224 : (shape == 1 ? 128 : (shape == 2 ? 140 : 168))) / 2 + 1;           // q0 / 2 <= wb / 2
    const int sizeL = lanes_h + tpad + q0h + 8;
    const int sizeS = q0h + tpad + lanes_h + 8;
    return (uint32_t)((sizeS + sizeL + 15) & ~15);
}

cudaError_t launch_linear(const LinearParams &p, bool wide, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(do_linear_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(do_linear_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        once.mark(dev);
    }
    if (wide) do_linear_kernel<32><<<grid, block, smem_bytes, stream>>>(p);
    else      do_linear_kernel<8><<<grid, block, smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

// worst-case staging bytes for strings of up to max_len elements with B diagonals per lane
uint32_t seq_cap_for(int max_len, int B)
{
    const int H = B / 2;
    const int tpad = (max_len + 1 + 15) & ~15;
    const int q0h = 16 * B + 1;                       // q0 / 2 <= wb / 2 <= 16 B
    const int sizeL = 32 * H + tpad + q0h + 8;
    const int sizeS = q0h + tpad + 32 * H + 8;
    return (uint32_t)((2 * sizeS + sizeL + 15) & ~15);
}

}  // namespace pcgdo
