// do_pack16.cuh — shared pieces of the packed 16-bit fill (two diagonals per register, VIADDMNMX.U16x2); the
// scheme is described in do_linear.cu.  Used by the dense-table kernel (do_linear.cu) and the per-pair memo-table
// kernel (do_explicit.cu).
#pragma once
#include "do_device.cuh"

namespace pcgdo {

constexpr unsigned kFull = 0xffffffffu;
// Iterations unrolled per chunk of the packed fill: 16 up to 8 diagonals per lane, 8 beyond — the unrolled body of the
// wider variants would otherwise be tens of KB and fall out of the instruction cache.  With 32 diagonals per lane the
// register windows rotate with period 16, so every 8-iteration chunk ends with a rotation by register moves (rot16).
__host__ __device__ constexpr int pack16_chunk(int B) { return B <= 8 ? 16 : 8; }
template <int H, int S>
__device__ __forceinline__ void rot16(uint32_t (&a)[H])
{
    uint32_t t[H];
#pragma unroll
    for (int x = 0; x < H; ++x) t[x] = a[(x + S) % H];
#pragma unroll
    for (int x = 0; x < H; ++x) a[x] = t[x];
}

constexpr uint32_t kTop16 = 0x7000u;
constexpr uint32_t kOverflow16 = 0xffffffffu;


template <int H>
__device__ __forceinline__ void store_words16(uint32_t *dst, const uint32_t (&dw)[H], int n)
{
    const int sh = 2 * (8 - n);           // left-align a partial group; the halves cannot collide (2n + sh = 16)
    if constexpr (H >= 4) {
#pragma unroll
        for (int m = 0; m < H; m += 4)
            *reinterpret_cast<uint4 *>(dst + m) = make_uint4(dw[m] << sh, dw[m + 1] << sh, dw[m + 2] << sh, dw[m + 3] << sh);
    } else {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(dw[0] << sh, dw[1] << sh);
    }
}


// Move the warp's finite maximum back to kTop16.  Returns false when the finite spread no longer fits the window.
template <int H>
__device__ __forceinline__ bool rebase16(uint32_t (&R)[H], int &total_shift, uint32_t spread_limit)
{
    uint32_t mx = 0u, mnv = 0xffffffffu;
#pragma unroll
    for (int m = 0; m < H; ++m) {
        mx = __vmaxu2(mx, R[m] ^ 0x80008000u);       // finite halves become >= 0x8000, unreachable ones < 0x8000
        mnv = __vminu2(mnv, R[m]);
    }
    uint32_t hi = max(mx & 0xffffu, mx >> 16), lo = min(mnv & 0xffffu, mnv >> 16);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        hi = max(hi, __shfl_xor_sync(kFull, hi, d));
        lo = min(lo, __shfl_xor_sync(kFull, lo, d));
    }
    if (hi < 0x8000u) return true;                   // no finite cell yet
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


// one lane per pair, packed layout
__device__ __forceinline__ uint32_t trace_lane16(const uint32_t *dirw, int B, int q0, int lg, int sh, uint32_t *ops)
{
    // B is a power of two (4 .. 32).  The walk is one long dependent chain per lane — and, with one lane per pair, an
    // instruction stream the warp issues in full however few of its lanes hold a pair (a launch with 2 pairs per warp
    // spent a third of its instructions here).  State is (s, q) = (i + j, diagonal); an ALIGN move stays on its diagonal
    // and, 7 times out of 8, inside its direction word: only the other moves recompute the word's address.  One move per
    // iteration for every lane, so the lanes of a warp stay in step.
    const int H = B >> 1, lh = __ffs(H) - 1;
    int s = (lg - 1) + (sh - 1), q = (sh - 1) - (lg - 1) + q0;
    uint32_t n = 0, acc = 0, w = 0;
    int sft = 16;
    while (s > 0) {
        if (sft == 16) {
            const int t = (s - (q & 1)) >> 1;
            // diagonal q = lane ln, slot mm of its B; slot mm lives in register mm % H, half mm / H
            const int half = (q >> lh) & 1, col = ((q >> (lh + 1)) << lh) | (q & (H - 1));     // col = ln * H + reg
            w = __ldcg(dirw + (((uint32_t)(t >> 3) << (lh + 5)) + (uint32_t)col)) >> (16 * half);
            sft = 2 * (7 - (t & 7));
        }
        const uint32_t code = (w >> sft) & 3u;
        acc |= code << (2 * (n & 15));
        ++n;
        if ((n & 15) == 0) { ops[(n >> 4) - 1] = acc; acc = 0; }
        if (code == 0u) { s -= 2; sft += 2; }
        else { s -= 1; q += (code == 1u) ? 1 : -1; sft = 16; }      // 1: consume longer (i--), 2: consume lesser (j--)
    }
    if (n & 15) ops[n >> 4] = acc;
    return n;
}

}  // namespace pcgdo
