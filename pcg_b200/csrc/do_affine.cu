// do_affine.cu — affine-gap pairwise direct optimization ("C-affine" dialect) for sm_100a.
//
// Replaces align2dAffine (EXT/c_alignment_interface.c:180-348): plane initialisation
// (EXT/alignCharacters.c:3124-3215), the banded 4-plane fill (:3483-3727 with FILL_EXTEND_HORIZONTAL /
// _VERTICAL / _BLOCK_DIAGONAL, FILL_CLOSE_BLOCK_DIAGONAL :2523-2848 and ASSIGN_MINIMUM :3218-3256) and the
// traceback state machine with its median rules (:2850-3010).  PARITY UNPINNED upstream (no golden vector;
// the substitution lookup at :3636 is undefined behaviour): the device table `aff_sub` is built on the host
// for either reading of that expression ("as coded on x86-64" or "patched"), the kernel is the same.
//
// Geometry.  Rows i = 1..S follow the SHORTER string, columns j = 1..L the longer one.  The reference's band
// (start_pos advancing after row 40, end_pos = max(40, L-S+8) growing by one per row) is the diagonal range
// d = j - i in [-40, E0] where d = -40 and column 0 hold its vertical-only border cells and d = E0 its
// all-infinite right border.  One warp owns one pair; lane k owns B consecutive diagonals and keeps the four
// plane values of each diagonal's latest cell in registers.  A ROW is processed per step (the warp is uniform
// in i, so every per-row quantity of the reference is a broadcast load): CB / EV / EB of row i depend on row
// i-1 only and are computed for all cells at once; EH, the only plane with an in-row dependency
// (EH[j] = min(EH[j-1] + he[j], CB[j-1] + go[j] + gr[j])), is a min-plus prefix scan over the row
// (lane-local composition + 5 shuffle steps).  Each cell's traceback information (which plane wins, with the
// reference's priority horizontal > align > vertical > diagonal; the next mode out of an align step,
// horizontal > diagonal > vertical > stay; the three END flags) is one byte; four rows make a word per
// diagonal, stored as B consecutive words per lane.  Trace runs one lane per pair after a warp has filled
// its hand-out of pairs; emit is warp-cooperative with coalesced stores.
#include "do_device.cuh"

namespace pcgdo {

constexpr unsigned kFullA = 0xffffffffu;
constexpr uint32_t kAInf = 1u << 28;          // "very large": finite values stay far below it (checked on the host)
constexpr uint32_t kACap = 2u << 28;          // clamp for sums of "very large" values
constexpr int kDB = -40;                      // left border diagonal (alignCharacters.c:3500 start_v = 40)

struct AffGeom { int S, L, E0, wq_need; };

__host__ __device__ inline AffGeom make_aff_geom(int S, int L)
{
    AffGeom g;
    g.S = S; g.L = L;
    int e = (L - S) + 8;
    if (e < 40) e = 40;
    if (e > L) e = L;
    g.E0 = e;
    g.wq_need = e - kDB + 1;                  // diagonals -40 .. E0
    return g;
}
__host__ __device__ inline int aff_choose_b(int wq_need)
{
    if (wq_need <= 128) return 4;
    if (wq_need <= 256) return 8;
    if (wq_need <= 512) return 16;
    if (wq_need <= 1024) return 32;
    return 0;
}
// wavefront fill: finer steps (6 and 12 diagonals per lane: a quarter fewer cell updates for bands just above 128 / 256)
__host__ __device__ inline int aff_choose_b_wave(int wq_need)
{
    if (wq_need <= 128) return 4;
    if (wq_need <= 192) return 6;
    if (wq_need <= 256) return 8;
    if (wq_need <= 384) return 12;
    if (wq_need <= 512) return 16;
    if (wq_need <= 1024) return 32;       // row-wise fill
    return 0;
}
// per-pair scratch (32-bit words): direction words | op stream | column params (2 words each) | row params
__host__ __device__ inline uint32_t aff_dir_words(const AffGeom &g, int B) { return (uint32_t)(g.S / 4 + 1) * 32u * (uint32_t)B; }
__host__ __device__ inline uint32_t aff_ops_words(const AffGeom &g) { return ((uint32_t)(g.S + g.L) / 16u + 4u) & ~1u; }
__host__ __device__ inline uint32_t aff_col_entries(const AffGeom &g, int B) { return (uint32_t)(g.L + 32 * B + 96); }
__host__ __device__ inline uint32_t aff_need_words(const AffGeom &g, int B)
{
    return (aff_dir_words(g, B) + aff_ops_words(g) + 2u * aff_col_entries(g, B) + 2u * (uint32_t)(g.S + 8) +
            (uint32_t)(g.L + 8) + 7u) & ~3u;
}
constexpr int kColPad = 48;                   // column-parameter entries in front of j = 0 (j runs down to 1 - 40 - ...)

struct AffParams {
    TcmDev    tcm;
    BatchDev  b;
    const uint16_t *aff_sub;   // [s' << (a-1) | l'] substitution costs between gap-stripped elements
    uint32_t  gap_open;
    uint32_t *arena;
    uint32_t  arena_words;
    int       pairs_per_warp;
    unsigned int *counter;
    unsigned int *overflow_count;
    int       wave;            // 1: wavefront fill (default), 0: the row-wise fill with its prefix scan
    unsigned int *class_count;     // [6] pairs per band class (filled by aff_classify_kernel)
    unsigned int *class_take;      // [6] hand-out counters
    uint32_t     *big_list;        // [6][n_pairs] pair indices per class
    uint32_t      k_one;           // 1, opaque to the compiler (keeps additions on the FMA pipe as IMAD)
};

// column entry: .x = gr (cost[gap][l_j]); .y = l'_j | lgap << 8 | go_is_open << 9 | he_has_open << 10
// row entry   : .x = ve;                  .y = ge | so_is_open << 16 | sgap << 17 | s'_i << 24

template <int B, bool EARLY>
__device__ __forceinline__ void aff_row(uint32_t (&CB)[B], uint32_t (&EB)[B], uint32_t (&EV)[B], uint32_t (&EH)[B],
                                        const uint32_t (&penup)[B], uint32_t (&fl)[B], const uint2 *colpar, uint2 rp,
                                        uint32_t GO, const uint16_t *sub, int sub_shift, int i, int d0, int lane)
{
    const uint32_t ve = rp.x, ge = rp.y & 0xffffu, so = ((rp.y >> 16) & 1u) ? GO : 0u, sgap = (rp.y >> 17) & 1u;
    const uint16_t *subrow = sub + ((rp.y >> 24) << sub_shift);
    // up neighbours: diagonal q+1 of the previous row; the block edge comes from the next lane
    const uint32_t nEV = __shfl_down_sync(kFullA, EV[0], 1), nCB = __shfl_down_sync(kFullA, CB[0], 1);
    uint32_t cbn[B], evn[B], ebn[B], ha[B], hb[B];
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int j = i + d0 + m;
        const uint2 cp = colpar[j];
        const uint32_t gr = cp.x, lq = cp.y & 0xffu, lgap = (cp.y >> 8) & 1u;
        const uint32_t go = ((cp.y >> 9) & 1u) ? GO : 0u, he = gr + (((cp.y >> 10) & 1u) ? go : 0u);
        uint32_t uEV = (m == B - 1) ? nEV : EV[m + 1 < B ? m + 1 : m];
        uint32_t uCB = (m == B - 1) ? nCB : CB[m + 1 < B ? m + 1 : m];
        const uint32_t pen = (EARLY && i == 1) ? 0u : penup[m];   // row 1 still sees the real row-0 cell there
        uEV |= pen;
        uCB |= pen;
        // extend vertical (FILL_EXTEND_VERTICAL)
        const uint32_t vx = uEV + ve, vo = uCB + so + ge;
        uint32_t ev = min(vx, vo);
        uint32_t endv = !(vx < vo);
        // extend block diagonal (FILL_EXTEND_BLOCK_DIAGONAL): both options add the same term
        uint32_t eb = (sgap & lgap) ? min(EB[m], CB[m]) : kAInf;
        const uint32_t endb = !(EB[m] < CB[m]);
        // close block diagonal (FILL_CLOSE_BLOCK_DIAGONAL)
        const uint32_t dg = subrow[lq];
        const uint32_t ca = CB[m] + dg;
        const uint32_t fv = EV[m] + dg + (sgap ? go : 0u);
        const uint32_t fh = EH[m] + dg + (lgap ? so : 0u);
        const uint32_t fd = EB[m] + dg + max(go, so);
        uint32_t cb = min(min(ca, fv), min(fh, fd));
        uint32_t an = 3u;                       // next mode out of an align step: H > D > V > stay
        an = (fv == cb) ? 2u : an;
        an = (fd == cb) ? 1u : an;
        an = (fh == cb) ? 0u : an;
        const bool border = EARLY ? (j == 0) : (m == 0 && lane == 0);   // vertical extension only (:3624-3634)
        const bool outside = EARLY && j < 0;
        if (border) { ev = vx; cb = kAInf; eb = kAInf; endv = 1u; }
        if (outside) { ev = kAInf; cb = kAInf; eb = kAInf; }
        cbn[m] = min(cb, kACap); evn[m] = min(ev, kACap); ebn[m] = eb;
        fl[m] = (an << 2) | (endv << 4) | (endb << 6);
        // horizontal extension as the min-plus map x -> min(x + a, b); b gets CB[j-1] of THIS row below
        ha[m] = (border || outside) ? 0u : he;
        hb[m] = go + gr;
    }
    const uint32_t lCB = __shfl_up_sync(kFullA, cbn[B - 1], 1);
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int j = i + d0 + m;
        const uint32_t left = (m == 0) ? (lane == 0 ? kAInf : lCB) : cbn[m > 0 ? m - 1 : 0];
        const bool dead = EARLY ? (j <= 0) : (m == 0 && lane == 0);
        hb[m] = dead ? kACap : min(left + hb[m], kACap);
    }
    uint32_t A = 0, Bv = kACap;                 // composition of this lane's maps, then exclusive scan over lanes
#pragma unroll
    for (int m = 0; m < B; ++m) { Bv = min(Bv + ha[m], hb[m]); A += ha[m]; }
    A = min(A, kAInf);
    Bv = min(Bv, kACap);
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const uint32_t A2 = __shfl_up_sync(kFullA, A, dlt), B2 = __shfl_up_sync(kFullA, Bv, dlt);
        if (lane >= dlt) { Bv = min(min(B2 + A, Bv), kACap); A = min(A2 + A, kAInf); }
    }
    uint32_t x = __shfl_up_sync(kFullA, Bv, 1);   // EH just left of this lane's first cell
    if (lane == 0) x = kACap;
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const uint32_t hx = x + ha[m], ho = hb[m];
        uint32_t eh = min(min(hx, ho), kACap);
        const uint32_t endh = !(hx < ho);
        x = eh;
        // ASSIGN_MINIMUM: which plane ends here, priority horizontal > align > vertical > diagonal
        const uint32_t fc = min(min(eh, evn[m]), min(ebn[m], cbn[m]));
        uint32_t td = 3u;
        td = (evn[m] == fc) ? 2u : td;
        td = (cbn[m] == fc) ? 1u : td;
        td = (eh == fc) ? 0u : td;
        fl[m] |= td | (endh << 5);
        CB[m] = cbn[m]; EV[m] = evn[m]; EB[m] = ebn[m]; EH[m] = eh;
    }
}

template <int B>
__device__ __forceinline__ void aff_store(uint32_t *dst, const uint32_t (&dw)[B])
{
#pragma unroll
    for (int m = 0; m < B; m += 4) *reinterpret_cast<uint4 *>(dst + m) = make_uint4(dw[m], dw[m + 1], dw[m + 2], dw[m + 3]);
}

// Fills one pair; returns FC at (S, L).  colpar points at the entry of column j = 0.
template <int B>
__device__ __noinline__ uint32_t aff_fill(const AffGeom g, uint32_t GO, const uint16_t *sub, int sub_shift,
                                          const uint2 *colpar, const uint2 *rowpar, const uint32_t *row0,
                                          uint32_t *dirw, int lane)
{
    uint32_t CB[B], EB[B], EV[B], EH[B], penup[B], fl[B], dw[B];
    const int d0 = kDB + lane * B;
    const int qE = g.E0 - kDB;                   // diagonal index of the right border
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int q = lane * B + m, d = kDB + q;
        penup[m] = (q + 1 >= qE) ? kAInf : 0u;   // the cell above the last in-band diagonal is "very large"
        // row 0 (:3152-3173): cell (0, d)
        CB[m] = EB[m] = EV[m] = EH[m] = kAInf;
        if (d == 0) { CB[m] = 0; EB[m] = 0; EV[m] = GO; EH[m] = GO; }
        else if (d > 0 && d <= g.L) { CB[m] = row0[d]; EH[m] = row0[d]; }
        dw[m] = 0;
    }
    uint32_t fc_last = 0;
    const int q_last = (g.L - g.S) - kDB, lane_last = q_last / B, m_last = q_last % B;
    int i = 1;
    const int early_end = g.S < 40 ? g.S : 40;
    for (; i <= early_end; ++i) {
        aff_row<B, true>(CB, EB, EV, EH, penup, fl, colpar, rowpar[i], GO, sub, sub_shift, i, d0, lane);
#pragma unroll
        for (int m = 0; m < B; ++m) dw[m] |= fl[m] << (8 * (i & 3));
        if ((i & 3) == 3) {
            aff_store<B>(dirw + (size_t)(i >> 2) * 32 * B + lane * B, dw);
#pragma unroll
            for (int m = 0; m < B; ++m) dw[m] = 0;
        }
    }
    for (; i <= g.S; ++i) {
        aff_row<B, false>(CB, EB, EV, EH, penup, fl, colpar, rowpar[i], GO, sub, sub_shift, i, d0, lane);
#pragma unroll
        for (int m = 0; m < B; ++m) dw[m] |= fl[m] << (8 * (i & 3));
        if ((i & 3) == 3) {
            aff_store<B>(dirw + (size_t)(i >> 2) * 32 * B + lane * B, dw);
#pragma unroll
            for (int m = 0; m < B; ++m) dw[m] = 0;
        }
    }
    if ((g.S & 3) != 3) aff_store<B>(dirw + (size_t)(g.S >> 2) * 32 * B + lane * B, dw);
    // FC of the last cell = min over the four planes there
    uint32_t v = 0;
#pragma unroll
    for (int m = 0; m < B; ++m)
        if (m == m_last) v = min(min(EH[m], EV[m]), min(EB[m], CB[m]));
    fc_last = __shfl_sync(kFullA, v, lane_last);
    return fc_last;
}

// One lane per pair.  Op codes: 0 vertical (consume shorter), 1 horizontal (consume longer), 2 block diagonal,
// 3 align.  Written in traceback order (last column first), 2 bits each.
__device__ __forceinline__ uint32_t aff_trace(const uint32_t *dirw, uint32_t wq, int S, int L, uint32_t *ops)
{
    int si = S, li = L;
    uint32_t n = 0, acc = 0;
    int mode = 0;                                  // 0 todo, 1 vertical, 2 horizontal, 3 diagonal, 4 align
    auto put = [&](uint32_t code) {
        acc |= code << (2 * (n & 15));
        ++n;
        if ((n & 15) == 0) { ops[(n >> 4) - 1] = acc; acc = 0; }
    };
    while (si != 0 && li != 0) {
        const int q = li - si - kDB;
        const uint32_t f = (__ldcg(dirw + (size_t)(si >> 2) * wq + q) >> (8 * (si & 3))) & 0xffu;
        if (mode == 0) {
            const uint32_t td = f & 3u;            // 0 H, 1 A, 2 V, 3 D
            mode = td == 0 ? 2 : td == 1 ? 4 : td == 2 ? 1 : 3;
        } else if (mode == 1) {
            if (f & 16u) mode = 0;
            put(0); --si;
        } else if (mode == 2) {
            if (f & 32u) mode = 0;
            put(1); --li;
        } else if (mode == 3) {
            if (f & 64u) mode = 0;
            put(2); --si; --li;
        } else {
            const uint32_t an = (f >> 2) & 3u;     // 0 H, 1 D, 2 V, 3 stay
            mode = an == 0 ? 2 : an == 1 ? 3 : an == 2 ? 1 : 4;
            put(3); --si; --li;
        }
    }
    while (si != 0) { put(0); --si; }
    while (li != 0) { put(1); --li; }
    if (n & 15) ops[n >> 4] = acc;
    return n;
}

// ---------------------------------------------------------------------------------------------
// Wavefront fill (default).  Same cells, same per-cell rules as aff_row above, but swept along ANTI-DIAGONALS like
// the linear kernel: lane k owns B consecutive diagonals q = d + 40, iteration t handles anti-diagonal s = 2t for
// even q and s = 2t + 1 for odd q, the four plane values of each diagonal's latest cell stay in registers and
// are updated in place.  The only in-row dependency of the reference (EH on the cell to its left) is then a
// neighbour of the PREVIOUS half step, so no prefix scan is needed; the up / left neighbours at the lane edges
// come from four shuffles per iteration.  All plane values are kept multiplied by 4 so that the two priority
// decisions of a cell are single mins over candidates that carry their priority in the low two bits
// (next mode out of an align step: horizontal 0 > diagonal 1 > vertical 2 > stay 3; plane that ends here:
// horizontal 0 > align 1 > vertical 2 > diagonal 3).  Row / column parameters are decoded once per lane and
// iteration into sliding register windows (a row or column is used by B cells of the lane).  The first
// iterations, in which some diagonal of the warp is still above row 1 or left of column 1, run a predicated
// variant (EARLY); row 0 lives in the registers' initial values.
// Direction bytes: word[(t >> 2) * Wq + q], byte t & 3 — same flag layout as the row-wise fill.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kInf4 = 1u << 29;          // "very large" x 4
constexpr uint32_t kCap4 = 1u << 30;          // clamp for sums of very large values

// parameter entries (global scratch, 16 bytes each), decoded into the windows below
//   row i   : .x = 4*ve + 1, .y = 4*(so + ge), .z = 4*so, .w = table row offset (bytes) | sgap << 31
//   column j: .x = 4*he + 1, .y = 4*(go + gr), .z = 4*go, .w = table column offset (bytes) | lgap << 31
//   (the + 1 marks "the run continues" in the candidate that extends it, see aff_cell)
template <int H>
struct AffWin { uint32_t a[H], b[H], c[H], m[H], o[H]; };

template <int H>
__device__ __forceinline__ void aff_win_load(AffWin<H> &w, int slot, const uint4 v, uint32_t obase)
{
    w.a[slot] = v.x; w.b[slot] = v.y; w.c[slot] = v.z;
    w.m[slot] = (uint32_t)((int32_t)v.w >> 31) & 0xfffffffcu;      // all-ones above the two tag bits, or zero
    w.o[slot] = (v.w & 0x7fffffffu) + obase;                       // rows carry the table's shared address
}

template <int H, int S>
__device__ __forceinline__ void aff_rot(uint32_t (&a)[H])
{
    uint32_t t[H];
#pragma unroll
    for (int x = 0; x < H; ++x) t[x] = a[(x + S) % H];
#pragma unroll
    for (int x = 0; x < H; ++x) a[x] = t[x];
}
template <int H, int S>
__device__ __forceinline__ void aff_win_rotate(AffWin<H> &w)
{
    aff_rot<H, S>(w.a); aff_rot<H, S>(w.b); aff_rot<H, S>(w.c); aff_rot<H, S>(w.m); aff_rot<H, S>(w.o);
}

__device__ __forceinline__ uint32_t aff_lds32(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// iterations unrolled per chunk of the wavefront fill (see aff_wave_chunk)
__host__ __device__ constexpr int aff_wave_chunk_len(int B) { return B >= 12 ? 2 : 4; }

// a * k + c on the FMA pipe (k is a run-time 1 or 4..64 the compiler cannot fold): the cell update is otherwise all
// add / min / logic, i.e. all ALU pipe, which ncu showed saturated (73 % busy, math-pipe throttle the top stall) while
// the FMA pipe idled at 10 %.  Every plain addition of the cell therefore goes through IMAD.
__device__ __forceinline__ uint32_t aff_imad(uint32_t a, uint32_t k, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(k), "r"(c));
    return d;
}

// One cell.  rs / cs: window slots of its row and column.  Returns the flag byte.
// Plane values are multiples of 4; every two-way decision is taken by ONE min over candidates that carry the decision
// in their low bits (the staged parameters ve and he arrive with +1 baked in: "this candidate extends the run"):
//   ev  = min(uEV + ve + 1, uCB + so + ge)      bit 0 set  <=> the vertical run continues   (endv = !bit0)
//   eh  = min(lEH + he + 1, lCB + go + gr)      bit 0 set  <=> the horizontal run continues (endh = !bit0)
//   eb  = min(EB + 1, CB)                       bit 0 set  <=> the diagonal run continues   (endb = !bit0)
//   cb  = min(fh, fd | 1, fv | 2, ca | 3)       low two bits = next mode out of an align step (H > D > V > stay)
//   fc  = min(eh, cb | 1, ev | 2, eb | 3)       low two bits = plane that ends here         (H > A > V > D)
// ties resolve exactly as the reference's strict comparisons do: the candidate with the smaller tag wins.
template <int H, bool EARLY>
__device__ __forceinline__ uint32_t aff_cell(uint32_t &CB, uint32_t &EB, uint32_t &EV, uint32_t &EH, uint32_t uEV,
                                             uint32_t uCB, uint32_t lEH, uint32_t lCB, uint32_t penup,
                                             const AffWin<H> &rw, int rs, const AffWin<H> &cw, int cs, uint32_t k1,
                                             bool border, int i, int j)
{
    const uint32_t ve1 = rw.a[rs], voa = rw.b[rs], so = rw.c[rs], sgM = rw.m[rs];
    const uint32_t he1 = cw.a[cs], hoa = cw.b[cs], go = cw.c[cs], lgM = cw.m[cs];
    const uint32_t pen = (EARLY && i == 1) ? 0u : penup;      // row 1 still sees the real row-0 cell above the band
    // extend vertical: the cell above the last in-band diagonal is "very large" (pen)
    const uint32_t t = aff_imad(uEV, k1, ve1);
    uint32_t evm = max(__viaddmin_u32(uCB, voa, t), pen);
    uint32_t ev = evm & ~3u;
    // extend block diagonal (both strings have a gap here)
    const uint32_t gm = sgM & lgM;
    const uint32_t ebm = __viaddmin_u32(EB, 1u, CB);
    uint32_t eb = (ebm & gm) | (kInf4 & ~gm);
    // close block diagonal: candidates carry the next mode in their low bits
    const uint32_t dg = aff_lds32(aff_imad(rw.o[rs], k1, cw.o[cs]));
    const uint32_t fh = aff_imad(lgM & so, k1, aff_imad(EH, k1, dg));
    const uint32_t fd = aff_imad((go | so) | 1u, k1, aff_imad(EB, k1, dg));
    const uint32_t fv = aff_imad((sgM & go) | 2u, k1, aff_imad(EV, k1, dg));
    const uint32_t ca = CB + dg + 3u;
    const uint32_t cbm = min(__vimin3_u32(fh, fd, fv), ca);
    uint32_t cb = cbm & ~3u;
    // extend horizontal: the left neighbour of this half step
    uint32_t ehm = __viaddmin_u32(lCB, hoa, aff_imad(lEH, k1, he1));
    uint32_t eh = ehm & ~3u;
    // (overridden cells: their markers are made to read "run ends here", evm = ev and ehm = eh)
    if (border) { ev = t & ~3u; evm = ev; cb = kInf4; eb = kInf4; eh = kCap4; ehm = eh; }     // d = -40 / column 0: vertical extension only
    if (EARLY && j < 0) { ev = kInf4; evm = ev; cb = kInf4; eb = kInf4; eh = kCap4; ehm = eh; }
    // no clamping: every candidate is ONE plane value plus a small term, so "very large" values only drift by
    // 4 * (a few costs) per step — they never wrap and never come near a finite value (checked on the host)
    // which plane ends here: the tags go on by addition (the values are clean multiples of 4), on the FMA pipe
    const uint32_t fcm = min(__vimin3_u32(eh, aff_imad(cb, k1, 1u), aff_imad(ev, k1, 2u)), eb | 3u);
    // flag byte td | an << 2 | endv << 4 | endh << 5 | endb << 6, end* = 1 - marker.  A marker is the difference between
    // a decision value and its cleaned copy (evm - ev, ehm - eh, cbm - cb for `an`), so the byte is assembled by
    // multiply-adds alone: 112 + td + 4 (cbm - cb) - 16 (evm - ev) - 32 (ehm - eh) - 64 (ebm & 1)
    uint32_t fl = aff_imad(cbm, 4u * k1, (fcm & 3u) + 112u);
    fl = aff_imad(cb, 0u - 4u * k1, fl);
    fl = aff_imad(ev, 16u * k1, fl);
    fl = aff_imad(evm, 0u - 16u * k1, fl);
    fl = aff_imad(eh, 32u * k1, fl);
    fl = aff_imad(ehm, 0u - 32u * k1, fl);
    fl = aff_imad(ebm & 1u, 0u - 64u * k1, fl);
    if (!EARLY || i >= 1) { CB = cb; EB = eb; EV = ev; EH = eh; }             // rows <= 0: keep the row-0 value
    return fl;
}

template <int B, bool EARLY, bool TAIL>
__device__ __forceinline__ void aff_wave_chunk(uint32_t (&CB)[B], uint32_t (&EB)[B], uint32_t (&EV)[B], uint32_t (&EH)[B],
                                               const uint32_t (&penup)[B], uint32_t (&dw)[B], AffWin<B / 2> &rw,
                                               AffWin<B / 2> &cw, uint32_t tab, uint32_t k1, const uint4 *&rptr,
                                               const uint4 *&cptr, int n_iter, int i_even0, int j_even0, int lane,
                                               uint32_t *&out, int word_pos)
{
    // iterations unrolled per chunk: the body must stay well inside the 32 KB instruction cache (ncu: 23 % of the wide
    // kernel's stall samples were "no instruction" with 4 iterations of 16 diagonals = 51 KB)
    constexpr int H = B / 2, CH = aff_wave_chunk_len(B);
#pragma unroll
    for (int r = 0; r < CH; ++r) {
        if (TAIL && r >= n_iter) break;
        const int tt = r % H;
        // byte of the direction word this iteration fills: compile-time when a chunk is a whole word (CH = 4)
        const uint32_t sh8 = (CH % 4 == 0) ? (1u << (8 * (r & 3))) : (1u << (8 * ((word_pos + r) & 3)));
        // ---- even diagonals (anti-diagonal 2t): m = 2u handles row i_even0 + r - u, column j_even0 + r + u
        const uint32_t lEH0 = __shfl_up_sync(kFullA, EH[B - 1], 1), lCB0 = __shfl_up_sync(kFullA, CB[B - 1], 1);
#pragma unroll
        for (int u = 0; u < H; ++u) {
            const int m = 2 * u;
            const uint32_t lEH = (u == 0) ? (lane == 0 ? kCap4 : lEH0) : EH[m > 0 ? m - 1 : 0];
            const uint32_t lCB = (u == 0) ? (lane == 0 ? kCap4 : lCB0) : CB[m > 0 ? m - 1 : 0];
            const int i = i_even0 + r - u, j = j_even0 + r + u;
            const bool border = (u == 0 && lane == 0) || (EARLY && j == 0);     // d = -40 and column 0
            const uint32_t fl = aff_cell<H, EARLY>(CB[m], EB[m], EV[m], EH[m], EV[m + 1], CB[m + 1], lEH, lCB, penup[m],
                                                   rw, (tt - u + 2 * H) % H, cw, (tt + u) % H, k1, border, i, j);
            dw[m] = aff_imad(fl, sh8, dw[m]);
        }
        aff_win_load<H>(cw, tt, __ldg(cptr + r), 0u);
        // ---- odd diagonals (anti-diagonal 2t + 1): m = 2u + 1 handles row i_even0 + r - u, column j_even0 + r + u + 1
        const uint32_t uEV0 = __shfl_down_sync(kFullA, EV[0], 1), uCB0 = __shfl_down_sync(kFullA, CB[0], 1);
#pragma unroll
        for (int u = 0; u < H; ++u) {
            const int m = 2 * u + 1;
            const uint32_t uEV = (u == H - 1) ? uEV0 : EV[m + 1 < B ? m + 1 : B - 1];
            const uint32_t uCB = (u == H - 1) ? uCB0 : CB[m + 1 < B ? m + 1 : B - 1];
            const int i = i_even0 + r - u, j = j_even0 + r + u + 1;
            const bool border = EARLY ? (j == 0) : false;
            const uint32_t fl = aff_cell<H, EARLY>(CB[m], EB[m], EV[m], EH[m], uEV, uCB, EH[m - 1], CB[m - 1], penup[m],
                                                   rw, (tt - u + 2 * H) % H, cw, (tt + u + 1) % H, k1, border, i, j);
            dw[m] = aff_imad(fl, sh8, dw[m]);
        }
        aff_win_load<H>(rw, (tt + 1) % H, __ldg(rptr + r), tab);
        if (((word_pos + r) & 3) == 3 || (TAIL && r == n_iter - 1)) {
            if constexpr (B % 4 == 0) {
#pragma unroll
                for (int m = 0; m < B; m += 4) *reinterpret_cast<uint4 *>(out + m) = make_uint4(dw[m], dw[m + 1], dw[m + 2], dw[m + 3]);
            } else {
#pragma unroll
                for (int m = 0; m < B; m += 2) *reinterpret_cast<uint2 *>(out + m) = make_uint2(dw[m], dw[m + 1]);
            }
#pragma unroll
            for (int m = 0; m < B; ++m) dw[m] = 0;
            out += 32 * B;
        }
    }
    rptr += CH;
    cptr += CH;
    // the window slots are addressed with compile-time indices that assume the chunk starts at rotation 0: when the
    // chunk length is not a multiple of the rotation period, rotate the windows by register moves
    if constexpr (CH % H != 0) {
        aff_win_rotate<H, CH % H>(rw);
        aff_win_rotate<H, CH % H>(cw);
    }
}

// rowp / colp point at the entries of row 0 / column 0 (padded on both sides).  Returns 4 * FC at (S, L).
template <int B>
__device__ __noinline__ uint32_t aff_wave_fill(const AffGeom g, uint32_t GO, uint32_t tab, uint32_t k1, const uint4 *rowp,
                                               const uint4 *colp, const uint32_t *row0, uint32_t *dirw, int lane)
{
    constexpr int H = B / 2, CH = aff_wave_chunk_len(B);
    uint32_t CB[B], EB[B], EV[B], EH[B], penup[B], dw[B];
    AffWin<H> rw, cw;
    const int qE = g.E0 - kDB;
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int q = lane * B + m, d = kDB + q;
        penup[m] = (q + 1 >= qE) ? kInf4 : 0u;
        CB[m] = EB[m] = EV[m] = EH[m] = kInf4;
        if (d == 0) { CB[m] = 0; EB[m] = 0; EV[m] = 4u * GO; EH[m] = 4u * GO; }
        else if (d > 0 && d <= g.L) { CB[m] = 4u * row0[d]; EH[m] = 4u * row0[d]; }
        dw[m] = 0;
    }
    // even diagonal m = 2u of this lane at iteration t: row t + 20 - q/2, column t - 20 + q/2 with q = lane*B + 2u
    const int I0 = 20 - lane * H, J0 = -20 + lane * H;
#pragma unroll
    for (int u = 0; u < H; ++u) {
        aff_win_load<H>(rw, (H - u) % H, __ldg(rowp + (I0 - u)), tab);
        aff_win_load<H>(cw, u, __ldg(colp + (J0 + u)), 0u);
    }
    const uint4 *rptr = rowp + (I0 + 1), *cptr = colp + (J0 + H);
    uint32_t *out = dirw + lane * B;
    const int T = (g.S + g.L) / 2 + 1;
    int T0 = qE / 2 - 17;                         // from here on every diagonal is at row >= 1 and column >= 1
    if (T0 < 22) T0 = 22;
    T0 = (T0 + CH - 1) / CH * CH;
    int t = 0;
    for (; t + CH <= T0 && t + CH <= T; t += CH)
        aff_wave_chunk<B, true, false>(CB, EB, EV, EH, penup, dw, rw, cw, tab, k1, rptr, cptr, CH, I0 + t, J0 + t, lane, out, t & 3);
    for (; t + CH <= T; t += CH)
        aff_wave_chunk<B, false, false>(CB, EB, EV, EH, penup, dw, rw, cw, tab, k1, rptr, cptr, CH, I0 + t, J0 + t, lane, out, t & 3);
    if (t < T) {    // the last cell must stay in its register: stop exactly after iteration T - 1
        aff_wave_chunk<B, true, true>(CB, EB, EV, EH, penup, dw, rw, cw, tab, k1, rptr, cptr, T - t, I0 + t, J0 + t, lane, out, t & 3);
    } else if (CH % 4 != 0 && (T & 3) != 0) {
        // chunks shorter than a direction word: the last word is only partly filled and has not been stored yet
        if constexpr (B % 4 == 0) {
#pragma unroll
            for (int m = 0; m < B; m += 4) *reinterpret_cast<uint4 *>(out + m) = make_uint4(dw[m], dw[m + 1], dw[m + 2], dw[m + 3]);
        } else {
#pragma unroll
            for (int m = 0; m < B; m += 2) *reinterpret_cast<uint2 *>(out + m) = make_uint2(dw[m], dw[m + 1]);
        }
    }
    const int q_last = (g.L - g.S) - kDB, lane_last = q_last / B, m_last = q_last % B;
    uint32_t v = 0;
#pragma unroll
    for (int m = 0; m < B; ++m)
        if (m == m_last) v = min(min(EH[m], EV[m]), min(EB[m], CB[m]));
    return __shfl_sync(kFullA, v, lane_last);
}

// per-pair scratch of the wavefront fill (32-bit words): direction words | op stream | row params | column params | row 0
__host__ __device__ inline uint32_t aff_wave_iters(const AffGeom &g, int B)
{
    const int ch = aff_wave_chunk_len(B);
    return (uint32_t)((((g.S + g.L) / 2 + 1) + ch - 1) / ch * ch);
}
__host__ __device__ inline uint32_t aff_wave_dir_words(const AffGeom &g, int B) { return (aff_wave_iters(g, B) / 4u + 1u) * 32u * (uint32_t)B; }
constexpr int kWavePad = 16 * 32 + 64;        // parameter entries in front of index 0 (rows run down to 20 - 31*16)
__host__ __device__ inline uint32_t aff_wave_par_entries(const AffGeom &g, int B)
{
    return (uint32_t)(kWavePad + aff_wave_iters(g, B) + 32 * B + 64);
}
__host__ __device__ inline uint32_t aff_wave_ops_words(const AffGeom &g) { return (aff_ops_words(g) + 3u) & ~3u; }   // keeps the uint4 arrays aligned
__host__ __device__ inline uint32_t aff_wave_need_words(const AffGeom &g, int B)
{
    return (aff_wave_dir_words(g, B) + aff_wave_ops_words(g) + 8u * aff_wave_par_entries(g, B) + (uint32_t)(g.L + 8) + 7u) & ~3u;
}

// One lane per pair, wavefront layout.
__device__ __forceinline__ uint32_t aff_trace_wave(const uint32_t *dirw, uint32_t wq, int S, int L, uint32_t *ops)
{
    int si = S, li = L;
    uint32_t n = 0, acc = 0;
    int mode = 0;
    auto put = [&](uint32_t code) {
        acc |= code << (2 * (n & 15));
        ++n;
        if ((n & 15) == 0) { ops[(n >> 4) - 1] = acc; acc = 0; }
    };
    while (si != 0 && li != 0) {
        const int q = li - si - kDB;
        const int t = (si + li - (q & 1)) >> 1;
        const uint32_t f = (__ldcg(dirw + (size_t)(t >> 2) * wq + q) >> (8 * (t & 3))) & 0xffu;
        if (mode == 0) {
            const uint32_t td = f & 3u;
            mode = td == 0 ? 2 : td == 1 ? 4 : td == 2 ? 1 : 3;
        } else if (mode == 1) {
            if (f & 16u) mode = 0;
            put(0); --si;
        } else if (mode == 2) {
            if (f & 32u) mode = 0;
            put(1); --li;
        } else if (mode == 3) {
            if (f & 64u) mode = 0;
            put(2); --si; --li;
        } else {
            const uint32_t an = (f >> 2) & 3u;
            mode = an == 0 ? 2 : an == 1 ? 3 : an == 2 ? 1 : 4;
            put(3); --si; --li;
        }
    }
    while (si != 0) { put(0); --si; }
    while (li != 0) { put(1); --li; }
    if (n & 15) ops[n >> 4] = acc;
    return n;
}

__device__ __forceinline__ void aff_emit(const TcmDev &tcm, const uint8_t *pL, const uint8_t *pS, const uint32_t *ops,
                                         uint32_t nal, bool first_long, uint8_t *og, uint8_t *o1, uint8_t *o2, int lane)
{
    uint32_t base_l = 0, base_s = 0;
    const uint32_t g = tcm.gap, amb = g - 1u;
    for (uint32_t c0 = 0; c0 < nal; c0 += 128) {
        const uint32_t c = c0 + 4u * lane;
        uint32_t code[4], nl = 0, ns = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const uint32_t cc = c + x;
            code[x] = 4u;
            if (cc < nal) {
                const uint32_t r = nal - 1u - cc;
                code[x] = (__ldcg(ops + (r >> 4)) >> (2 * (r & 15))) & 3u;
            }
            nl += (code[x] == 1u || code[x] == 2u || code[x] == 3u);
            ns += (code[x] == 0u || code[x] == 2u || code[x] == 3u);
        }
        const uint32_t packed = nl | (ns << 16);
        uint32_t incl = packed;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(kFullA, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t excl = incl - packed;
        uint32_t il = base_l + (excl & 0xffffu), is = base_s + (excl >> 16);
        uint32_t wg = 0, ws = 0, wl = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            uint32_t se = g, le = g, md = 0;
            if (code[x] != 4u) {
                if (code[x] != 0u) { le = __ldg(pL + il); ++il; }
                if (code[x] != 1u) { se = __ldg(pS + is); ++is; }
                if (code[x] == 0u)      md = (se & g) ? g : (se | g);          // vertical   (:2904-2916)
                else if (code[x] == 1u) md = (le & g) ? g : (le | g);          // horizontal (:2917-2931)
                else if (code[x] == 2u) md = g;                                // block diagonal
                else md = __ldg(tcm.median + (((se & amb) << tcm.a) | (le & amb)));
            } else { se = 0; le = 0; }
            wg |= md << (8 * x);
            ws |= se << (8 * x);
            wl |= le << (8 * x);
        }
        if (c < nal) {
            if (og) *reinterpret_cast<uint32_t *>(og + c) = wg;
            if (o1) *reinterpret_cast<uint32_t *>(o1 + c) = first_long ? wl : ws;   // slots keep their own string
            if (o2) *reinterpret_cast<uint32_t *>(o2 + c) = first_long ? ws : wl;
        }
        const uint32_t tot = __shfl_sync(kFullA, incl, 31);
        base_l += tot & 0xffffu;
        base_s += tot >> 16;
    }
}

extern __shared__ __align__(16) char aff_smem[];

__device__ __forceinline__ unsigned long long aff_band_cells(const AffGeom &ag)
{
    unsigned long long c = 0;
    for (int i = 1; i <= ag.S; ++i) {
        int lo = i > 40 ? i - 39 : 1, hi = ag.E0 + i - 1;
        if (hi > ag.L) hi = ag.L;
        if (hi >= lo) c += (unsigned long long)(hi - lo + 1);
    }
    return c;
}

// Stage the parameters of one pair in its scratch slot and run the wavefront fill.  Returns FC at (S, L).
template <bool WIDE>
__device__ __forceinline__ uint32_t aff_pair_wave(const TcmDev &tcm, const AffGeom &ag, int B, uint32_t GO, uint32_t tab_lane,
                                                  uint32_t ent_shift, const uint8_t *pL, const uint8_t *pS, uint32_t *slot,
                                                  int lane, uint32_t k1)
{
    const uint32_t g = tcm.gap;
    uint32_t *dirw = slot;
    uint32_t *ops = dirw + aff_wave_dir_words(ag, B);
    const uint32_t npar = aff_wave_par_entries(ag, B);
    uint4 *rowbase = reinterpret_cast<uint4 *>(ops + aff_wave_ops_words(ag));
    uint4 *colbase = rowbase + npar;
    uint4 *rowp = rowbase + kWavePad, *colp = colbase + kWavePad;       // entries of row 0 / column 0
    uint32_t *row0 = reinterpret_cast<uint32_t *>(colbase + npar);      // L + 1 words: EH = CB of row 0
    uint32_t carry = GO;
    for (int base_e = 0; base_e < (int)npar; base_e += 32) {
        const int e = base_e + lane, j = e - kWavePad, i = j;
        uint4 cv = make_uint4(0, 0, 0, 0), rv = make_uint4(0, 0, 0, 0);
        uint32_t gr = 0;
        if (j >= 1 && j <= ag.L) {                                      // column parameters (:3585-3595)
            const uint32_t lj = __ldg(pL + j - 1), lp = j == 1 ? g : __ldg(pL + j - 2);
            gr = tcm.gapc[lj];
            const uint32_t go = !(!(g & lp) && (g & lj)) ? GO : 0u;
            const uint32_t he = gr + ((j > 1 && (lp & g) && !(lj & g)) ? go : 0u);
            cv = make_uint4(4u * he + 1u, 4u * (go + gr), 4u * go, ((lj & (g - 1u)) << ent_shift) | ((lj & g) ? 0x80000000u : 0u));
        }
        if (i >= 1 && i <= ag.S) {                                      // row parameters (:3614-3625)
            const uint32_t si = __ldg(pS + i - 1), sp = i == 1 ? g : __ldg(pS + i - 2);
            const uint32_t ge = tcm.cgap[si];
            const uint32_t so = !(!(g & sp) && (g & si)) ? GO : 0u;
            const uint32_t ve = (i > 1 && (sp & g) && !(si & g)) ? so + ge : ge;
            rv = make_uint4(4u * ve + 1u, 4u * (so + ge), 4u * so,
                            (((si & (g - 1u)) << (tcm.a - 1)) << ent_shift) | ((si & g) ? 0x80000000u : 0u));
        }
        if (e < (int)npar) { colbase[e] = cv; rowbase[e] = rv; }
        uint32_t sfx = gr;                                              // inclusive prefix of gr for row 0
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(kFullA, sfx, d);
            if (lane >= d) sfx += o;
        }
        if (j >= 1 && j <= ag.L) row0[j] = carry + sfx;
        carry += __shfl_sync(kFullA, sfx, 31);
    }
    __syncwarp();
    uint32_t fc4;
    if constexpr (WIDE) {
        if (B == 12) fc4 = aff_wave_fill<12>(ag, GO, tab_lane, k1, rowp, colp, row0, dirw, lane);
        else         fc4 = aff_wave_fill<16>(ag, GO, tab_lane, k1, rowp, colp, row0, dirw, lane);
    } else {
        if (B == 4)      fc4 = aff_wave_fill<4>(ag, GO, tab_lane, k1, rowp, colp, row0, dirw, lane);
        else if (B == 6) fc4 = aff_wave_fill<6>(ag, GO, tab_lane, k1, rowp, colp, row0, dirw, lane);
        else             fc4 = aff_wave_fill<8>(ag, GO, tab_lane, k1, rowp, colp, row0, dirw, lane);
    }
    return fc4 >> 2;
}

// The row-wise fill of one pair (column / row parameters in the scratch slot, prefix scan per row).
template <bool WIDE>
__device__ __forceinline__ uint32_t aff_pair_rowwise(const TcmDev &tcm, const AffGeom &ag, int B, uint32_t GO,
                                                     const uint16_t *s_sub, const uint8_t *pL, const uint8_t *pS,
                                                     uint32_t *slot, int lane)
{
    const uint32_t g = tcm.gap;
    uint32_t *dirw = slot;
    uint32_t *ops = dirw + aff_dir_words(ag, B);
    uint2 *colbase = reinterpret_cast<uint2 *>(ops + aff_ops_words(ag));
    uint2 *colpar = colbase + kColPad;                            // entry of column j = 0
    uint2 *rowpar = colbase + aff_col_entries(ag, B);
    uint32_t *row0 = reinterpret_cast<uint32_t *>(rowpar + ag.S + 8);   // L + 1 words: EH = CB of row 0
    // ---- per-column parameters (:3585-3595) and row 0 prefix; per-row parameters (:3614-3625)
    const int ncol = (int)aff_col_entries(ag, B);
    uint32_t carry = GO;
    for (int base_j = -kColPad; base_j < ncol - kColPad; base_j += 32) {
        const int j = base_j + lane;
        uint32_t gr = 0, y = 0;
        if (j >= 0 && j <= ag.L) {
            const uint32_t lj = j == 0 ? g : __ldg(pL + j - 1);
            const uint32_t lp = j <= 1 ? g : __ldg(pL + j - 2);       // l_{j-1}; l_0 = gap
            if (j >= 1) {
                gr = tcm.gapc[lj];
                const uint32_t go_open = !(!(g & lp) && (g & lj));
                uint32_t he_open = ((lp & g) && !(lj & g)) ? go_open : 0u;
                if (j == 1) he_open = 0u;                             // he[1] = gr[1]
                y = (lj & (g - 1u)) | (((lj & g) ? 1u : 0u) << 8) | (go_open << 9) | (he_open << 10);
            }
        }
        if (j < ncol - kColPad) colpar[j] = make_uint2(gr, y);
        uint32_t sfx = gr;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(kFullA, sfx, d);
            if (lane >= d) sfx += o;
        }
        if (j >= 1 && j <= ag.L) row0[j] = carry + sfx;
        carry += __shfl_sync(kFullA, sfx, 31);
    }
    for (int i = 1 + lane; i <= ag.S; i += 32) {
        const uint32_t si = __ldg(pS + i - 1), sp = i == 1 ? g : __ldg(pS + i - 2);
        const uint32_t ge = tcm.cgap[si];
        const uint32_t so_open = !(!(g & sp) && (g & si));
        const uint32_t so = so_open ? GO : 0u;
        const uint32_t ve = (i > 1 && (sp & g) && !(si & g)) ? so + ge : ge;
        rowpar[i] = make_uint2(ve, ge | (so_open << 16) | (((si & g) ? 1u : 0u) << 17) | ((si & (g - 1u)) << 24));
    }
    __syncwarp();
    if (ag.S <= 0) return ag.L > 0 ? row0[ag.L] : 0u;
    if constexpr (WIDE) {
        return aff_fill<32>(ag, GO, s_sub, tcm.a - 1, colpar, rowpar, row0, dirw, lane);
    } else {
        switch (B) {
            case 4:  return aff_fill<4>(ag, GO, s_sub, tcm.a - 1, colpar, rowpar, row0, dirw, lane);
            case 8:  return aff_fill<8>(ag, GO, s_sub, tcm.a - 1, colpar, rowpar, row0, dirw, lane);
            case 16: return aff_fill<16>(ag, GO, s_sub, tcm.a - 1, colpar, rowpar, row0, dirw, lane);
            default: return aff_fill<32>(ag, GO, s_sub, tcm.a - 1, colpar, rowpar, row0, dirw, lane);
        }
    }
}

// Pairs are sorted into per-band-class lists first: the fill variants are long unrolled loops, and warps of one SM
// running different variants at the same time thrash the instruction cache (measured: 1.7x slower on config 3).
// class 0..5 = 4, 6, 8 | 12, 16, 32 diagonals per lane (wavefront chooser) or 4, -, 8 | -, 16, 32 (row-wise chooser).
__host__ __device__ inline int aff_class_of(int B) { return B == 4 ? 0 : B == 6 ? 1 : B == 8 ? 2 : B == 12 ? 3 : B == 16 ? 4 : 5; }

__global__ void aff_classify_kernel(const AffParams p)
{
    const BatchDev &b = p.b;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < b.n_pairs; k += gridDim.x * blockDim.x) {
        const uint32_t n1 = b.len1[k], n2 = b.len2[k];
        const AffGeom ag = make_aff_geom((int)(n1 >= n2 ? n2 : n1), (int)(n1 >= n2 ? n1 : n2));
        const int B = p.wave ? aff_choose_b_wave(ag.wq_need) : aff_choose_b(ag.wq_need);
        if (!B) { atomicAdd(p.overflow_count, 1u); continue; }
        const int c = aff_class_of(B);
        p.big_list[(size_t)c * b.n_pairs + atomicAdd(p.class_count + c, 1u)] = k;
    }
}

// WIDE = false: 512 threads (128 registers); takes pairs straight from the batch.  With the wavefront fill it keeps
//               the pairs whose band fits 8 diagonals per lane and queues the others on big_list.
// WIDE = true : 256 threads (255 registers: 16 diagonals per lane with four planes and two parameter windows
//               each); takes its pairs from big_list.  Bands beyond 512 diagonals use the row-wise fill.
// Shared memory: [both substitution tables | per warp: flags[32], nal[32], off[32], pair[32]].
template <bool WIDE>
__global__ void __launch_bounds__(WIDE ? 256 : 512, 1) do_affine_kernel(const AffParams p)
{
    const TcmDev &tcm = p.tcm;
    const BatchDev &b = p.b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int hdim = 1 << (tcm.a - 1);
    const bool repl = hdim <= 16;
    // wavefront fill: 32-bit entries 4 * cost, for small alphabets replicated 32 times with copy k wholly in bank k;
    // row-wise fill: the plain 16-bit table
    const size_t wave_bytes = p.wave ? (size_t)hdim * hdim * (repl ? 128 : 4) : 0;
    uint16_t *s_sub = reinterpret_cast<uint16_t *>(aff_smem + wave_bytes);
    const size_t tab_bytes = wave_bytes + (((size_t)hdim * hdim * 2 + 15) & ~(size_t)15);
    uint32_t *s_meta = reinterpret_cast<uint32_t *>(aff_smem + tab_bytes) + warp * 128;
    uint32_t *s_flags = s_meta, *s_nal = s_meta + 32, *s_off = s_meta + 64, *s_pair = s_meta + 96;
    if (p.wave) {
        uint32_t *t32 = reinterpret_cast<uint32_t *>(aff_smem);
        if (repl) {
            for (int x = threadIdx.x; x < hdim * hdim * 32; x += blockDim.x) t32[x] = 4u * p.aff_sub[x >> 5];
        } else {
            for (int x = threadIdx.x; x < hdim * hdim; x += blockDim.x) t32[x] = 4u * p.aff_sub[x];
        }
    }
    for (int x = threadIdx.x; x < hdim * hdim; x += blockDim.x) s_sub[x] = p.aff_sub[x];
    __syncthreads();
    const uint32_t tab_lane = (uint32_t)__cvta_generic_to_shared(aff_smem) + (repl ? 4u * lane : 0u);
    const uint32_t ent_shift = repl ? 7 : 2;      // bytes per table entry: 128 (replicated) or 4

    const uint32_t GO = p.gap_open;
    const int G = p.pairs_per_warp;
    uint32_t *arena = p.arena + ((size_t)blockIdx.x * nwarps + warp) * (size_t)p.arena_words;
    const bool want_out = b.out_gapped || b.out_slot1 || b.out_slot2 || b.out_len;
    // classes this kernel owns: wavefront mode 0..2 (narrow) / 3..5 (wide); row-wise mode: everything in the narrow one
    int cls = WIDE ? 3 : 0;
    const int cls_end = (WIDE || !p.wave) ? 6 : 3;

    // pairs are heavy (a million cells): take them one at a time, trace up to G of them at once
    uint32_t gpos = 0, gend = 0;
    bool exhausted = false;
    for (;;) {
        int cnt = 0;
        uint32_t cur = 0;
        for (;;) {
            if (cnt == G) break;
            if (gpos == gend) {
                if (exhausted) break;
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(p.class_take + cls, 1u);
                base = __shfl_sync(kFullA, base, 0);
                if (base >= p.class_count[cls]) {                        // this class is handed out: on to the next
                    if (++cls >= cls_end) { exhausted = true; break; }
                    continue;
                }
                gpos = base;
                gend = base + 1;
            }
            const uint32_t k = p.big_list[(size_t)cls * b.n_pairs + gpos];
            const uint32_t n1 = b.len1[k], n2 = b.len2[k];
            const bool fl = n1 >= n2;                                    // c_alignment_interface.c:212
            const uint8_t *pL = b.elems + (fl ? b.off1[k] : b.off2[k]), *pS = b.elems + (fl ? b.off2[k] : b.off1[k]);
            const AffGeom ag = make_aff_geom((int)(fl ? n2 : n1), (int)(fl ? n1 : n2));
            const int B = p.wave ? aff_choose_b_wave(ag.wq_need) : aff_choose_b(ag.wq_need);
            const bool wave = p.wave && ag.S > 0 && B <= 16;
            const uint32_t need = !B ? 0u : wave ? aff_wave_need_words(ag, B) : aff_need_words(ag, B);
            if (!B || need > p.arena_words) {
                if (lane == 0) atomicAdd(p.overflow_count, 1u);
                ++gpos;
                continue;
            }
            if (cur + need > p.arena_words) break;
            uint32_t *slot = arena + cur;
            const uint32_t fc = wave ? aff_pair_wave<WIDE>(tcm, ag, B, GO, tab_lane, ent_shift, pL, pS, slot, lane, p.k_one)
                                     : aff_pair_rowwise<WIDE>(tcm, ag, B, GO, s_sub, pL, pS, slot, lane);
            if (lane == 0) {
                s_flags[cnt] = (fl ? 1u : 0u) | (wave ? 4u : 0u) | ((uint32_t)B << 8);
                s_off[cnt] = cur;
                s_pair[cnt] = k;
                b.cost[k] = (int32_t)fc;
                if (b.cells) b.cells[k] = aff_band_cells(ag);
            }
            cur += need;
            ++cnt;
            ++gpos;
            __syncwarp();
        }
        __syncwarp();
        if (cnt && want_out) {
            if (lane < cnt) {
                const uint32_t k = s_pair[lane];
                const uint32_t n1 = b.len1[k], n2 = b.len2[k];
                const bool fl = s_flags[lane] & 1u;
                const AffGeom ag = make_aff_geom((int)(fl ? n2 : n1), (int)(fl ? n1 : n2));
                const int B = (int)(s_flags[lane] >> 8);
                uint32_t *slot = arena + s_off[lane];
                const uint32_t n = (s_flags[lane] & 4u)
                                       ? aff_trace_wave(slot, 32u * B, ag.S, ag.L, slot + aff_wave_dir_words(ag, B))
                                       : aff_trace(slot, 32u * B, ag.S, ag.L, slot + aff_dir_words(ag, B));
                s_nal[lane] = n;
                if (b.out_len) b.out_len[k] = n;
            }
            __syncwarp();
            if (b.out_gapped || b.out_slot1 || b.out_slot2) {
                for (int pe = 0; pe < cnt; ++pe) {
                    const uint32_t k = s_pair[pe];
                    const uint32_t n1 = b.len1[k], n2 = b.len2[k];
                    const bool fl = s_flags[pe] & 1u;
                    const AffGeom ag = make_aff_geom((int)(fl ? n2 : n1), (int)(fl ? n1 : n2));
                    const int B = (int)(s_flags[pe] >> 8);
                    const uint32_t *ops = arena + s_off[pe] + ((s_flags[pe] & 4u) ? aff_wave_dir_words(ag, B) : aff_dir_words(ag, B));
                    const uint64_t oo = b.out_off[k];
                    aff_emit(tcm, b.elems + (fl ? b.off1[k] : b.off2[k]), b.elems + (fl ? b.off2[k] : b.off1[k]), ops,
                             s_nal[pe], fl, b.out_gapped ? b.out_gapped + oo : nullptr,
                             b.out_slot1 ? b.out_slot1 + oo : nullptr, b.out_slot2 ? b.out_slot2 + oo : nullptr, lane);
                }
            }
            __syncwarp();
        }
        if (exhausted && gpos == gend) break;
    }
}

uint32_t aff_words_for(uint32_t n1, uint32_t n2)
{
    const AffGeom g = make_aff_geom((int)(n1 < n2 ? n1 : n2), (int)(n1 < n2 ? n2 : n1));
    const int B = aff_choose_b(g.wq_need), Bw = aff_choose_b_wave(g.wq_need);
    if (!B || !Bw) return 0u;
    const uint32_t w1 = aff_need_words(g, B), w2 = aff_wave_need_words(g, Bw);
    return w1 > w2 ? w1 : w2;
}

cudaError_t launch_affine_classify(const AffParams &p, int grid, cudaStream_t stream)
{
    aff_classify_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_affine(const AffParams &p, bool wide, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
    static DeviceOnce once;
    int dev = 0;
    if (once.pending(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(do_affine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(do_affine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        once.mark(dev);
    }
    if (wide) do_affine_kernel<true><<<grid, block, smem_bytes, stream>>>(p);
    else      do_affine_kernel<false><<<grid, block, smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace pcgdo
