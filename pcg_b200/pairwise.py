"""Host-side mirror of the reference's pairwise-DO operator interface, over the C-ABI CUDA library.

Reference names kept (DO = lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization):

* ``align2d_batch``            — many raw ``align2d(...,0,1,0)`` problems in one launch
                                 (EXT/c_alignment_interface.c:16-177 semantics per pair).
* ``foreign_pairwise_do``      — ``foreignPairwiseDO`` (DO/Pairwise/FFI.hsc:179-260): ungap, measure /
  ``foreign_pairwise_do_batch``  swap, align on the GPU, re-insert gaps, swap back.

A dynamic character is ``None`` (Missing) or an ``(n, 3)`` uint8 array of (median, left, right)
ambiguity sets (DYN/Internal.hs:86-89), gap = ``1 << (a - 1)``.

The DP, traceback and median all run in CUDA; the O(m+n) gap bookkeeping of the Haskell wrapper is
numpy here (it is Haskell in the reference), never an alignment.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .backend import Batch, Batch32, Context, Tcm


def _rup4(x):
    return (x + 3) & ~3


@dataclass
class PackedPairs:
    """Host-side packing of a batch in the layout ``pcgdo_batch`` wants."""
    n: int
    elems: np.ndarray      # uint8, all strings back to back
    off1: np.ndarray       # uint64
    off2: np.ndarray
    len1: np.ndarray       # uint32
    len2: np.ndarray
    out_off: np.ndarray    # uint64, multiples of 4
    out_bytes: int

    @staticmethod
    def from_lists(first: Sequence[np.ndarray], second: Sequence[np.ndarray]) -> "PackedPairs":
        n = len(first)
        assert n == len(second)
        len1 = np.fromiter((len(x) for x in first), dtype=np.uint32, count=n)
        len2 = np.fromiter((len(x) for x in second), dtype=np.uint32, count=n)
        tot = len1.astype(np.uint64) + len2.astype(np.uint64)
        starts = np.concatenate([[0], np.cumsum(tot)]).astype(np.uint64)
        off1 = starts[:-1].copy()
        off2 = off1 + len1
        elems = np.zeros(int(starts[-1]) + 16, np.uint8)
        for k in range(n):
            elems[int(off1[k]):int(off1[k]) + int(len1[k])] = first[k]
            elems[int(off2[k]):int(off2[k]) + int(len2[k])] = second[k]
        cap = _rup4(tot.astype(np.int64))
        ostarts = np.concatenate([[0], np.cumsum(cap)]).astype(np.uint64)
        return PackedPairs(n, elems, off1, off2, len1, len2, ostarts[:-1].copy(), int(ostarts[-1]) + 16)

    @staticmethod
    def uniform(strings: np.ndarray) -> "PackedPairs":
        """``strings``: (2n, L) uint8; pair k = rows (2k, 2k+1). No per-pair Python work."""
        two_n, L = strings.shape
        n = two_n // 2
        elems = np.ascontiguousarray(strings).reshape(-1)
        k = np.arange(n, dtype=np.uint64)
        off1 = k * np.uint64(2 * L)
        off2 = off1 + np.uint64(L)
        ln = np.full(n, L, np.uint32)
        cap = _rup4(2 * L)
        return PackedPairs(n, elems, off1, off2, ln, ln.copy(), k * np.uint64(cap), int(n * cap) + 16)


@dataclass
class BatchResult:
    cost: np.ndarray                   # int32
    out_len: Optional[np.ndarray]      # uint32
    gapped: Optional[np.ndarray]       # uint8 flat, indexed by PackedPairs.out_off
    slot1: Optional[np.ndarray]
    slot2: Optional[np.ndarray]
    cells: Optional[np.ndarray]        # uint64
    out_off: Optional[np.ndarray] = None   # compact transfer: where each pair's outputs actually are

    def pair(self, pp: PackedPairs, k: int):
        o, n = int((pp.out_off if self.out_off is None else self.out_off)[k]), int(self.out_len[k])
        return int(self.cost[k]), self.gapped[o:o + n], self.slot1[o:o + n], self.slot2[o:o + n]


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def align2d_batch(ctx: Context, tcm: Tcm, pp: PackedPairs, want_strings: bool = True,
                  want_cells: bool = False, compact: bool = False, pinned: bool = False) -> BatchResult:
    """Host buffers in / host buffers out through ``pcgdo_align2d_batch_host``.  ``compact``: only the bytes the
    alignments produced come back over PCIe (``pcgdo_batch.out_off_actual``); ``BatchResult.pair`` follows the offsets.
    ``pinned``: the three output arrays are pinned host memory — with ``compact`` the kernel then writes them directly."""
    n = pp.n
    cost = np.zeros(n, np.int32)
    out_len = np.zeros(n, np.uint32) if want_strings else None
    g = s1 = s2 = None
    if want_strings:
        g, s1, s2 = ((ctx.pinned_empty(pp.out_bytes) if pinned else np.zeros(pp.out_bytes, np.uint8)) for _ in range(3))
    cells = np.zeros(n, np.uint64) if want_cells else None
    actual = np.zeros(n, np.uint64) if (compact and want_strings) else None
    b = Batch(n, _p(pp.elems), pp.elems.nbytes, _p(pp.off1), _p(pp.off2), _p(pp.len1), _p(pp.len2),
              _p(pp.out_off), pp.out_bytes, _p(g), _p(s1), _p(s2), _p(out_len), _p(cost), _p(cells), _p(actual))
    if getattr(tcm, "gap_open", 0):       # getAlignmentStrategy (FFI.hsc:287-290): affine iff cost_model_type says so
        ctx.check(ctx.lib.pcgdo_align2d_affine_batch_host(ctx.h, tcm.h, C.byref(b)), "pcgdo_align2d_affine_batch_host")
    else:
        ctx.check(ctx.lib.pcgdo_align2d_batch_host(ctx.h, tcm.h, C.byref(b)), "pcgdo_align2d_batch_host")
    return BatchResult(cost, out_len, g, s1, s2, cells, actual)


# ------------------------------------------------------------------------------------------------
# Haskell wrapper
# ------------------------------------------------------------------------------------------------

def delete_gaps(c: np.ndarray, gap: int):
    """``deleteGaps`` (DYN/Internal.hs:135-185): (gaps-before-each-ungapped-index [n+1], ungapped)."""
    keep = c[:, 0] != gap
    idx = np.flatnonzero(keep)
    bounds = np.concatenate([[-1], idx, [len(c)]])
    gaps = (np.diff(bounds) - 1).astype(np.int64)
    return gaps, c[keep]


def _cmp_seq(x: np.ndarray, y: np.ndarray) -> int:
    d = np.flatnonzero(x != y)
    if len(d) == 0:
        return 0
    return -1 if x[d[0]] < y[d[0]] else 1


def measure_characters(x: np.ndarray, y: np.ndarray) -> int:
    """``measureCharacters`` ordering (DO/Pairwise/Internal.hs:307-337): -1 / 0 / 1."""
    if len(x) != len(y):
        return -1 if len(x) < len(y) else 1
    r = _cmp_seq(x[:, 0], y[:, 0])
    if r:
        return r
    return _cmp_seq(x.reshape(-1), y.reshape(-1))


def insert_gaps(gl: np.ndarray, gr: np.ndarray, aligned: np.ndarray, gap: int) -> np.ndarray:
    """``insertGaps`` (DYN/Internal.hs:189-241), vectorised: the aligned triple consuming lesser index p
    is preceded by gl[p] ``(gap,gap,0)`` then — once the longer index it sits at is reached — the pending
    ``(gap,0,gap)`` runs; positions follow from the running consumed counts."""
    if gl.sum() == 0 and gr.sum() == 0:
        return aligned
    n = len(aligned)
    # consumed counts BEFORE each aligned triple
    lp = np.concatenate([[0], np.cumsum(aligned[:, 1] != 0)])
    rp = np.concatenate([[0], np.cumsum(aligned[:, 2] != 0)])
    out = []
    gl = gl.copy()
    gr = gr.copy()
    ins = np.array([gap, gap, 0], np.uint8)
    dele = np.array([gap, 0, gap], np.uint8)
    for k in range(n + 1):
        a, b = int(lp[k]), int(rp[k])
        # gaps pending at the current (lp, rp) are emitted, left first, before the next aligned triple
        if a < len(gl) and gl[a]:
            out.append(np.tile(ins, (int(gl[a]), 1)))
            gl[a] = 0
        if b < len(gr) and gr[b]:
            out.append(np.tile(dele, (int(gr[b]), 1)))
            gr[b] = 0
        if k < n:
            out.append(aligned[k:k + 1])
    return np.concatenate(out, axis=0) if out else aligned


def foreign_pairwise_do_batch(ctx: Context, tcm: Tcm, chars1: Sequence[Optional[np.ndarray]],
                              chars2: Sequence[Optional[np.ndarray]], median_of_gap=None, _align=None
                              ) -> List[Tuple[int, Optional[np.ndarray]]]:
    """``foreignPairwiseDO`` over many pairs; one GPU launch for all non-trivial alignments."""
    n = len(chars1)
    gap = tcm.gap
    if median_of_gap is None:
        median_of_gap = getattr(tcm, "median_of_gap", None)
    res: List[Optional[Tuple[int, Optional[np.ndarray]]]] = [None] * n
    work = []
    firsts, seconds = [], []
    for k in range(n):
        c1, c2 = chars1[k], chars2[k]
        if c1 is None or c2 is None:                     # handleMissingCharacter
            res[k] = (0, c2 if (c1 is None and c2 is not None) else c1)
            continue
        c1 = np.asarray(c1, np.uint8).reshape(-1, 3)
        c2 = np.asarray(c2, np.uint8).reshape(-1, 3)
        g1, u1 = delete_gaps(c1, gap)
        g2, u2 = delete_gaps(c2, gap)
        order = measure_characters(u1, u2) or measure_characters(c1, c2)
        swapped = order > 0
        lesser, longer, gl, gr = (u2, u1, g2, g1) if swapped else (u1, u2, g1, g2)
        if len(lesser) == 0:
            if len(longer) == 0:
                al = np.zeros((0, 3), np.uint8)
                out = insert_gaps(gl, gr, al, gap)
                res[k] = (0, None if len(out) == 0 else (out[:, [0, 2, 1]] if swapped else out))
                continue
            # every element of the longer string is deleted against a gap (FFI.hsc:240-248);
            # medians come from the same dense table the device uses
            if median_of_gap is None:
                raise ValueError("median_of_gap table needed for the one-empty case")
            m = longer[:, 0]
            al = np.stack([median_of_gap[m], np.zeros_like(m), m], axis=1).astype(np.uint8)
            out = insert_gaps(gl, gr, al, gap)
            res[k] = (0, out[:, [0, 2, 1]] if swapped else out)
            continue
        work.append((k, swapped, gl, gr))
        firsts.append(np.ascontiguousarray(longer[:, 0]))      # longer is sent FIRST (FFI.hsc:261-268)
        seconds.append(np.ascontiguousarray(lesser[:, 0]))
    if work:
        pp = PackedPairs.from_lists(firsts, seconds)
        r = (_align or align2d_batch)(ctx, tcm, pp)     # _align: test hook for the host logic only
        for w, (k, swapped, gl, gr) in enumerate(work):
            cost, g, s1, s2 = r.pair(pp, w)
            al = np.stack([g, s1, s2], axis=1)
            out = insert_gaps(gl, gr, al, gap)
            res[k] = (cost, out[:, [0, 2, 1]] if swapped else out)
    return res  # type: ignore[return-value]


def foreign_pairwise_do(ctx: Context, tcm: Tcm, c1, c2, median_of_gap=None):
    return foreign_pairwise_do_batch(ctx, tcm, [c1], [c2], median_of_gap)[0]


# ------------------------------------------------------------------------------------------------
# Large alphabets: the Haskell dialects over the discrete metric (reference dispatch: DO/Internal.hs:67-83)
# ------------------------------------------------------------------------------------------------
S1_FULL, S1_UKKONEN, S2_UKKONEN, S2_FULL = 0, 1, 2, 3      # naiveDO / ukkonenDO / unboxedUkkonenFullSpaceDO / unboxedFullMatrixDO


@dataclass
class MetricResult:
    cost: np.ndarray
    out_len: np.ndarray
    median: np.ndarray          # uint32 flat, indexed by out_off
    a1: np.ndarray
    a2: np.ndarray
    cells: np.ndarray
    final_offset: np.ndarray
    out_off: np.ndarray

    def pair(self, k: int):
        o, n = int(self.out_off[k]), int(self.out_len[k])
        return int(self.cost[k]), self.median[o:o + n], self.a1[o:o + n], self.a2[o:o + n]


def metric_do_batch(ctx: Context, alphabet, dialect: int, first: Sequence[np.ndarray],
                    second: Sequence[np.ndarray]) -> MetricResult:
    """``naiveDO`` / ``ukkonenDO`` / ``unboxedUkkonenFullSpaceDO`` on median strings (uint32 ambiguity sets), one
    launch for the whole batch.  ``alphabet``: an int selects the discrete metric over that many symbols
    (``discreteMetricPairwiseLogic``); a ``Memo`` selects its overlap function — the explicit matrix (the reference's
    memoized ``getMedianAndCost2D`` path, ``extractPairwiseTransitionCostMatrix``, Bio/Metadata/Dynamic/Internal.hs:352-374)
    or the closed form PCG's TCM diagnosis picks.  Beyond 32 symbols strings are (n, W) uint32 arrays, W = ceil(a / 32),
    and so are the three outputs of ``MetricResult.pair``."""
    n = len(first)
    a = alphabet if isinstance(alphabet, int) else alphabet.a
    W = (a + 31) // 32
    first = [np.asarray(x, np.uint32).reshape(-1, W) for x in first]
    second = [np.asarray(x, np.uint32).reshape(-1, W) for x in second]
    len1 = np.fromiter((len(x) for x in first), dtype=np.uint32, count=n)
    len2 = np.fromiter((len(x) for x in second), dtype=np.uint32, count=n)
    tot = len1.astype(np.uint64) + len2.astype(np.uint64)
    starts = np.concatenate([[0], np.cumsum(tot)]).astype(np.uint64)
    off1 = starts[:-1].copy()
    off2 = off1 + len1
    elems = np.zeros((int(starts[-1]) + 4, W), np.uint32)
    for k in range(n):
        elems[int(off1[k]):int(off1[k]) + int(len1[k])] = first[k]
        elems[int(off2[k]):int(off2[k]) + int(len2[k])] = second[k]
    out_off = off1.copy()
    cnt = int(starts[-1]) + 4
    med, a1, a2 = (np.zeros((cnt, W), np.uint32) for _ in range(3))
    cost = np.zeros(n, np.int32)
    out_len = np.zeros(n, np.uint32)
    cells = np.zeros(n, np.uint64)
    fo = np.zeros(n, np.int32)
    b = Batch32(n, _p(elems), elems.size, _p(off1), _p(off2), _p(len1), _p(len2), _p(out_off), cnt * W,
                _p(med), _p(a1), _p(a2), _p(out_len), _p(cost), _p(cells), _p(fo))
    if isinstance(alphabet, int):
        ctx.check(ctx.lib.pcgdo_metric_batch_host(ctx.h, alphabet, dialect, C.byref(b)), "pcgdo_metric_batch_host")
    else:
        ctx.check(ctx.lib.pcgdo_explicit_batch_host(ctx.h, alphabet.h, dialect, C.byref(b)), "pcgdo_explicit_batch_host")
    if W == 1:
        med, a1, a2 = med.reshape(-1), a1.reshape(-1), a2.reshape(-1)
    return MetricResult(cost, out_len, med, a1, a2, cells, fo, out_off)
