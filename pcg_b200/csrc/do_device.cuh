// do_device.cuh — shared device-side types for the DO kernels (sm_100a).
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

namespace pcgdo {

// cudaFuncSetAttribute applies to the CURRENT device only: with several GPUs driven from one process (do_multi.cu, one
// host thread per device) every device needs its own call.  One bit per device ordinal.
struct DeviceOnce {
    std::atomic<unsigned long long> mask[4] = {};
    bool pending(int *dev)
    {
        if (cudaGetDevice(dev) != cudaSuccess) *dev = 0;
        return !((mask[(*dev >> 6) & 3].load() >> (*dev & 63)) & 1ull);
    }
    void mark(int dev) { mask[(dev >> 6) & 3].fetch_or(1ull << (dev & 63)); }
};

constexpr uint32_t kInfW   = 0x80000001u;  // "infinite" cell in stored form (bit 31 = out of band / unreachable)
constexpr uint32_t kBias4  = 1u << 30;     // 4 * BIAS; finite cells live in [0, 2^31)
constexpr int      kBigSub = 1 << 14;      // 4*sub'-1 entry for the padding element 0

// Device copy of one dense TCM.  Everything the kernels need, already in the layouts they read.
struct TcmDev {
    int            a;          // alphabet size (incl. gap), <= 8
    int            repl;       // 1: sub4 is the 32-way bank-replicated 32x32 table (a <= 5), 0: plain [l << a | s]
    uint32_t       gap;
    uint32_t       sub4_bytes; // bytes of `sub4` to stage in shared memory
    int32_t        sub4_gapgap;
    const int16_t *sub4;       // 4*(cost[l][s] - cost[l][gap] - cost[gap][s]) - 1 ; row/col 0 = kBigSub
    const int16_t *sub4_nat;   // the same table in plain [l << a | s] layout (general-geometry kernel)
    const uint8_t *median;     // [dim*dim]
    const uint32_t *cgap;      // [dim] cost[x][gap]   (indel of consuming x from the longer string)
    const uint32_t *gapc;      // [dim] cost[gap][y]
    // packed 16-bit fill (two diagonals per register, VIADDMNMX.U16x2): usable when every finite table entry lies
    // in [-s_minus, s_plus] with the margins do_linear.cu's rebase16 needs; spread16 = largest max - min of the
    // finite cell values of one anti-diagonal sweep that the 15-bit window can hold
    int            pack_ok;
    uint32_t       spread16;
};

struct BatchDev {
    uint32_t        n_pairs;
    const uint8_t  *elems;
    const uint64_t *off1, *off2;
    const uint32_t *len1, *len2;
    const uint64_t *out_off;
    uint8_t        *out_gapped, *out_slot1, *out_slot2;
    uint32_t       *out_len;
    int32_t        *cost;
    uint64_t       *cells;
    // post-order mode: the gapped median with its gap elements dropped (= the node's ungapped median
    // string, what the parent's alignment consumes), at out_uoff[k], at most ucap bytes, length to out_ulen
    uint8_t        *out_ungapped;
    const uint64_t *out_uoff;
    uint32_t       *out_ulen;
    uint32_t        ucap;
    // packed outputs (the host entry point's compact transfer): instead of out_off[k], a pair's three strings go to
    // offset atomicAdd(pack_cursor, round16(length)) of the out arrays, and out_off_actual[k] = pack_base + that offset
    // (out_off_actual[k] = ~0 and *pack_overflow += 1 when the range would end beyond pack_limit: the strings of that
    // pair are not written).  pack_gran = granularity - 1: 15 for a device staging slot, 3 when the out arrays are the
    // caller's pinned host buffers written over PCIe (never overflows the reserved capacity).
    unsigned long long *pack_cursor;
    uint64_t           *out_off_actual;
    uint64_t            pack_base, pack_limit;
    uint32_t            pack_gran;
    unsigned int       *pack_overflow;
    // added to the pair indices written to the failure list (a pipelined chunk reports batch-wide indices)
    uint32_t            pair_base;
};

__device__ __forceinline__ void pack_claim(const BatchDev &b, uint32_t k, uint32_t n)
{
    const unsigned long long sz = ((unsigned long long)n + b.pack_gran) & ~(unsigned long long)b.pack_gran;
    const unsigned long long oo = atomicAdd(b.pack_cursor, sz);
    if (oo + sz > b.pack_limit) {
        b.out_off_actual[k] = ~0ull;
        atomicAdd(b.pack_overflow, 1u);
    } else {
        b.out_off_actual[k] = b.pack_base + oo;
    }
}

struct LinearParams {
    TcmDev    tcm;
    BatchDev  b;
    uint32_t *arena;        // [n_warps_total][arena_words]: direction words + op streams of the
    uint32_t  arena_words;  //   pairs a warp has filled but not yet traced
    uint32_t  quad_arena_words;   // the same for a warp of the four-pairs-per-warp kernel (fewer warps per CTA share the arena)
    uint32_t  seq_cap;      // u16 elements of sequence staging per warp
    int       pairs_per_warp;   // pairs a warp fills before it traces them, one lane per pair (1..32)
    uint32_t  grab;         // pairs taken from the work counter at a time
    unsigned int *counter;  // work counter (pairs handed out); counter[2] = hand-out counter of the wide kernel
    unsigned int *overflow_count;
    uint32_t     *overflow_list;   // pair indices the fast path could not hold (first overflow_cap of them)
    uint32_t      overflow_cap;
    unsigned int *big_count;       // pairs whose band needs more than 8 diagonals per lane:
    uint32_t     *big_list;        //   queued by the narrow kernel, consumed by the wide one
    unsigned int *single_count;    // multi-pair kernels: the pairs the launch did not take (bit 31: 32-bit fill) — input of
    uint32_t     *single_list;     //   the next kernel; for the narrow kernel NULL = it reads the batch itself
    const unsigned int *multi_in_count;   // multi-pair kernels: the list to work from (NULL = the batch) and
    const uint32_t     *multi_in_list;    //   the hand-out counter of this launch
    unsigned int *multi_counter;
    // run-time "constants" (1, 4, -1, table row stride) kept opaque to ptxas so that the table-address add
    // and the direction accumulation are emitted as IMAD (FMA pipe) instead of IADD3/LEA/SHF (ALU pipe)
    uint32_t  k_four, k_minus1, l_stride, k_64k;
    // experiment switches of the four-pairs-per-warp kernel: bit 0 = stagger the warps' first batches; and, when not
    // NULL, per-phase cycle sums over all warps: [stage, fill, trace, emit]
    uint32_t  quad_flags;
    unsigned long long *phase_cycles;
};

// Geometry of one pair, identical on host and device (plain IEEE double multiplies only).
// lg / sh include the leading gap.  Mirrors the band selection of the reference
// (EXT/c_alignment_interface.c:101-109, EXT/alignCharacters.c:1163-1374) as ONE diagonal range
// d = j - i in [dlo, dhi]: outside cells are "infinite", which reproduces the reference's
// edge cells (left-most band cell without INSERT, right-most without DELETE) exactly.
struct Geom {
    int lg, sh;
    int dlo, dhi;
    int pad;        // 1 or 2 masked diagonals in front of the band: q = 0 is always masked, q0 is even
    int q0;         // diagonal index of d = 0
    int wb;         // diagonals incl. pad; the kernel also keeps q = 32*B - 1 masked, so it needs wb + 1 <= 32*B
    int T;          // iterations (two anti-diagonals each)
    int mode;       // 0 full matrix, 2 / 3 = the reference's banded cases
};

__host__ __device__ inline Geom make_geom(int lg, int sh)
{
    Geom g;
    g.lg = lg; g.sh = sh;
    const int diff = lg - sh;
    const int lim = (int)(.1 * (double)lg);
    const int delta = diff < lim ? lim / 2 : 2;
    int W = 50 + delta, H = diff + 50 + delta;
    if (W > sh) W = sh;
    if (H > lg) H = lg;
    int mode;
    if ((double)lg >= 1.5 * (double)sh) mode = 0;
    else if (2 * H < lg)                 mode = 2;
    else if (8 >= lg - H)                mode = 0;
    else                                 mode = 3;
    if (mode == 0)      { g.dlo = -(lg - 1); g.dhi = sh - 1; }
    else if (mode == 2) { g.dlo = -(H - 1);  g.dhi = W - 1; }
    else                { g.dlo = -H;        g.dhi = W - 1; }
    g.pad = 1 + ((1 - g.dlo) & 1);
    g.q0 = -g.dlo + g.pad;
    g.wb = g.dhi - g.dlo + 1 + g.pad;
    const int q_last = (sh - 1) - (lg - 1) + g.q0;
    g.T = ((lg - 1) + (sh - 1) - (q_last & 1)) / 2 + 1;
    g.mode = mode;
    return g;
}

// DP cells of the reference's band in rows i0, i0+step, ... (sum over lanes / the whole range = total).
__host__ __device__ inline unsigned long long band_cells(const Geom &g, int i0, int step)
{
    unsigned long long c = 0;
    for (int i = i0; i < g.lg; i += step) {
        int lo = i + g.dlo; if (lo < 0) lo = 0;
        int hi = i + g.dhi; if (hi > g.sh - 1) hi = g.sh - 1;
        if (hi >= lo) c += (unsigned long long)(hi - lo + 1);
    }
    return c;
}

}  // namespace pcgdo
