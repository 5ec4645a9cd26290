"""Executable numpy model of the B200 fill/trace/emit formulation (tests only, not product).

It mirrors, value for value, what pcg_b200/csrc/do_kernels.cu does on the device so the maths
(band -> diagonal range, bias/priority encoding, in-place parity update, direction-word layout,
traceback addressing) can be checked against the oracle on a CPU:

  * diagonals q = (j - i) + q0, q0 even, owned statically; anti-diagonal s = i + j;
    iteration t handles s = 2t (even q) then s = 2t + 1 (odd q), updating w[q] in place;
  * w = 4*(v' + BIAS) | code-bits, v'(i,j) = v(i,j) - C(i) - G(j)  (row / column indel prefix sums
    folded out so `up` needs no add and `left` only the +1 priority bump);
    candidates: diag = w[q] + (4*sub' - 1)  (code 0 = ALIGN), up = w[q+1] (code 1 = DELETE),
    left = w[q-1] + 1 (code 2 = INSERT); stored back as (min & ~3) | 1 | pen;
  * direction words: word[(t >> 4) * Wq + q] gets code << 2*(15 - (t & 15)); diagonals q = 0 and
    q = Wq - 1 are always masked so the block-edge shuffles of lanes 0 / 31 need no special case.
"""
import numpy as np

BIAS4 = 1 << 30          # 4 * BIAS
INF = 0x80000001
BIG = 1 << 14            # sub4 for padded (0) elements


def band_diagonals(lg, sh):
    """(dlo, dhi, cells) of align2d's band as a diagonal range d = j - i (lg/sh include the gap)."""
    diff = lg - sh
    lim = int(.1 * lg)
    delta = lim // 2 if diff < lim else 2
    W = min(sh, 50 + delta)
    H = min(lg, diff + 50 + delta)
    if lg >= 1.5 * sh:
        return -(lg - 1), sh - 1
    if 2 * H < lg:
        return -(H - 1), W - 1
    if 8 >= lg - H:
        return -(lg - 1), sh - 1
    return -H, W - 1


def fill(cost, a, lng, shr, lanes_b=None):
    """lng/shr include the leading gap. Returns (cost, words, geom)."""
    lg, sh = len(lng), len(shr)
    gap = 1 << (a - 1)
    dlo, dhi = band_diagonals(lg, sh)
    pad = 1 + ((1 - dlo) & 1)                 # q = 0 always masked, q0 even
    q0 = -dlo + pad
    wb = dhi - dlo + 1 + pad
    b = lanes_b or 2
    wq = -(-(wb + 1) // (32 * b)) * 32 * b    # the last diagonal of the warp stays masked too
    cost = cost.astype(np.int64)
    cg = cost[:, gap]                 # c(x) = cost[x][gap]
    gc = cost[gap, :]                 # g(y) = cost[gap][y]
    sub4 = 4 * (cost - cg[:, None] - gc[None, :]) - 1
    sub4[0, :] = BIG
    sub4[:, 0] = BIG
    Ctot = int(cg[lng[1:]].sum())     # C(0) = G(0) = 0: the leading gaps are never "consumed"
    Gtot = int(gc[shr[1:]].sum())
    q_last = (sh - 1) - (lg - 1) + q0
    T = ((lg - 1) + (sh - 1) - (q_last & 1)) // 2 + 1
    nchunk = (T + 15) // 16
    words = np.zeros((nchunk, wq), np.uint32)
    w = np.full(wq + 2, INF, np.int64)        # w[1 + q]; w[0] and w[wq+1] are the INF fences
    pen = np.zeros(wq, np.int64)
    pen[:pad] = 0x80000000
    pen[wb:] = 0x80000000
    w[1 + q0] = BIAS4 - int(sub4[gap, gap])
    L = np.zeros(lg + 4 * wq + 64, np.int64)
    S = np.zeros(sh + 4 * wq + 64, np.int64)
    offL = offS = 2 * wq + 32
    L[offL:offL + lg] = lng
    S[offS:offS + sh] = shr
    qs = np.arange(wq)
    final = None
    for t in range(T):
        for par in (0, 1):
            q = qs[par::2]
            d = q - q0
            s = 2 * t + par
            i = (s - d) // 2
            j = (s + d) // 2
            sb = sub4[L[offL + i], S[offS + j]]
            dg = w[1 + q] + sb
            up = w[2 + q]
            lf = w[q] + 1
            mn = np.minimum(np.minimum(dg, up), lf)
            assert (mn >= 0).all() and (mn < (1 << 32)).all()
            words[t >> 4, q] |= ((mn & 3) << (2 * (15 - (t & 15)))).astype(np.uint32)
            w[1 + q] = (mn & ~3) | 1 | pen[q]
        if t == T - 1:
            final = int(w[1 + q_last])
    v = (final >> 2) - (BIAS4 >> 2) + Ctot + Gtot
    return v, words, dict(q0=q0, wq=wq, T=T, dlo=dlo, dhi=dhi)


def trace(words, geom, lg, sh):
    """Returns op codes in alignment order (0 ALIGN, 1 DELETE = consume longer, 2 INSERT)."""
    q0, wq = geom["q0"], geom["wq"]
    i, j = lg - 1, sh - 1
    ops = []
    while i > 0 or j > 0:
        q = j - i + q0
        t = (i + j - (q & 1)) >> 1
        code = (int(words[t >> 4, q]) >> (2 * (15 - (t & 15)))) & 3
        ops.append(code)
        if code == 0:
            i -= 1; j -= 1
        elif code == 1:
            i -= 1
        else:
            assert code == 2
            j -= 1
    return ops[::-1]


def emit(median, a, lng, shr, ops):
    gap = 1 << (a - 1)
    ops = np.asarray(ops)
    adv_l = ops != 2
    adv_s = ops != 1
    il = np.cumsum(adv_l)
    js = np.cumsum(adv_s)
    y = np.where(adv_l, np.asarray(lng)[il], 0)
    x = np.where(adv_s, np.asarray(shr)[js], 0)
    med = median[np.where(x == 0, gap, x), np.where(y == 0, gap, y)]
    return med.astype(np.uint32), x.astype(np.uint32), y.astype(np.uint32)


def align2d(cost, median, a, in1, in2, lanes_b=None):
    """Same contract as oracle.align2d: (cost, gapped, slot1, slot2)."""
    in1 = np.asarray(in1, np.int64); in2 = np.asarray(in2, np.int64)
    first_long = True
    if len(in1) < len(in2):
        first_long = False
    elif len(in1) == len(in2):
        nz = np.nonzero(in1 != in2)[0]
        if len(nz):
            first_long = bool(in1[nz[0]] > in2[nz[0]])
    gap = 1 << (a - 1)
    lng = np.concatenate([[gap], in1 if first_long else in2])
    shr = np.concatenate([[gap], in2 if first_long else in1])
    c, words, geom = fill(cost, a, lng, shr, lanes_b)
    ops = trace(words, geom, len(lng), len(shr))
    med, x, y = emit(median, a, lng, shr, ops)
    return (c, med, x, y) if first_long else (c, med, y, x)
