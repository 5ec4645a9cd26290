// do_peak.cu — INT32 issue-rate micro-benchmark: the roofline denominator for the DP fill.
//
// The fill kernel is bound by integer instruction issue (add / min / logic on the ALU pipe, IMAD on
// the FMA pipe), not by HBM or tensor throughput, so bench.py measures this peak live on the same GPU
// at the clocks it actually sustains: 8 independent chains per thread of single-instruction integer
// operations (no fusable pairs, so operations == instructions).  mode 0: min + xor + add; mode 1: fused
// add-min (VIADDMNMX, ALU pipe) + multiply-add (IMAD, FMA pipe) — the two instructions the cell update is
// made of, one per pipe: the most any integer kernel can issue (measured 36.5 T lane-instr/s = 125.6 per
// clock per SM at 1965 MHz; either pipe alone tops out at 64 per clock per SM).
#include "do_device.cuh"

namespace pcgdo {

template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, int iters, uint32_t seed, uint32_t kk)
{
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 9u + i; b[i] = seed ^ (blockIdx.x * 131u + threadIdx.x * 7u + i * 77u); }
    const uint32_t k1 = seed | 1u, k2 = (seed >> 3) | 5u;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {       // keep the body small enough for the instruction cache
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 2) {
                    // the same pair with ONE run-time constant: ptxas's register assignment for mode 1 puts two source
                    // operands of its VIADDMNMX in one register bank (28.9 T measured); this form reaches 36.5 T
                    a[i] = __viaddmin_u32(a[i], kk, b[i]);
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(b[i]) : "r"(b[i]), "r"(kk), "r"(a[i]));
                } else if (MODE == 1) {
                    a[i] = __viaddmin_u32(a[i], k1, b[i]);                                             // ALU pipe (VIADDMNMX)
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(b[i]) : "r"(b[i]), "r"(k2), "r"(a[i]));    // FMA pipe (IMAD)
                } else {
                    a[i] = min(a[i], b[i]);                // VIMNMX
                    b[i] = (b[i] ^ k1) + a[i];             // LOP3 + add (ptxas picks IMAD.IADD): 2 instructions
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] ^ b[i];
    if (s == 0x12345678u) out[0] = s;
}

// Returns integer lane-instructions launched.
double launch_int_peak(int mode, int grid, int iters, uint32_t *scratch, cudaStream_t st)
{
    if (mode == 2)      int_peak_kernel<2><<<grid, 256, 0, st>>>(scratch, iters, 12345u, 3u);
    else if (mode == 1) int_peak_kernel<1><<<grid, 256, 0, st>>>(scratch, iters, 12345u, 3u);
    else                int_peak_kernel<0><<<grid, 256, 0, st>>>(scratch, iters, 12345u, 3u);
    return (double)grid * 256.0 * (double)iters * 8.0 * 8.0 * (mode >= 1 ? 2.0 : 3.0);   // SASS-checked counts
}

}  // namespace pcgdo
