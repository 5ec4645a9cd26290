"""ctypes loaders for the CPU oracle (TEST INFRASTRUCTURE ONLY — never imported by pcg_b200).

Two checkers live here:

* ``Oracle``      — ``oracle/liboracle_do.so``, our plain-C restatement (``do_oracle.c``).
* ``Reference``   — ``oracle/_ref/libpcgdo_ref.so``, the UNMODIFIED reference C aligner compiled
                    from ``/root/reference/lib/core/ffi/external-direct-optimization`` by
                    ``oracle/Makefile`` (target ``ref``).  Driven exactly like the Haskell binding
                    does (``DO/Pairwise/FFI.hsc:272-290, 474-515``): malloc'ed ``alignIO_t`` buffers
                    of capacity ``len1+len2+2`` with the payload right-aligned.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle_do.so")
REF_SO = os.path.join(_HERE, "_ref", "libpcgdo_ref.so")
REF_MEMO_SO = os.path.join(_HERE, "_ref", "libtcmmemo_ref.so")
REF_PATCHED_SO = os.path.join(_HERE, "_ref", "libpcgdo_ref_patched.so")

_u32p = C.POINTER(C.c_uint32)


def build(ref: bool = True) -> None:
    """Compile the restatement (always) and, if the reference tree is present, ``_ref``."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if ref and os.path.isdir("/root/reference") and not (os.path.exists(REF_SO) and os.path.exists(REF_PATCHED_SO)):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def _as_u32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint32))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_u32p)


class Oracle:
    """The C restatement (``do_oracle.c``)."""

    def __init__(self, sigma, gap_open: int = 0):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.oracle_tcm_new.restype = C.c_void_p
        L.oracle_tcm_new.argtypes = [_u32p, C.c_int, C.c_uint32]
        L.oracle_tcm_free.argtypes = [C.c_void_p]
        L.oracle_align2d.restype = C.c_int64
        L.oracle_align2d.argtypes = [C.c_void_p, _u32p, C.c_size_t, _u32p, C.c_size_t,
                                     _u32p, _u32p, _u32p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        L.oracle_foreign_pairwise_do.restype = C.c_int64
        L.oracle_foreign_pairwise_do.argtypes = (
            [C.c_void_p] + [_u32p, _u32p, _u32p, C.c_size_t, C.c_int] * 2
            + [_u32p, _u32p, _u32p, C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_uint64)])
        L.oracle_band_rows.restype = C.c_uint64
        L.oracle_band_rows.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(C.c_int),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)]
        sigma = _as_u32(sigma)
        self.a = int(sigma.shape[0])
        assert sigma.shape == (self.a, self.a)
        self.gap = 1 << (self.a - 1)
        self.sigma = sigma
        self.t = L.oracle_tcm_new(_ptr(sigma), self.a, gap_open)

        class _T(C.Structure):
            _fields_ = [("a", C.c_int), ("dim", C.c_uint32), ("gap", C.c_uint32), ("gap_open", C.c_uint32),
                        ("cost", _u32p), ("median", _u32p)]
        tt = C.cast(self.t, C.POINTER(_T)).contents
        dim = 1 << self.a
        self.cost = np.ctypeslib.as_array(tt.cost, shape=(dim, dim)).copy()
        self.median = np.ctypeslib.as_array(tt.median, shape=(dim, dim)).copy()

    def __del__(self):
        try:
            self.lib.oracle_tcm_free(self.t)
        except Exception:
            pass

    def band_cells(self, lg: int, sh: int):
        """(#cells, mode) for strings of lg/sh elements INCLUDING the leading gap."""
        mode = C.c_int(0)
        n = self.lib.oracle_band_rows(lg, sh, C.byref(mode), None, None, None)
        return int(n), int(mode.value)

    def align2d(self, in1, in2):
        """Raw ``align2d(...,0,1,0)``. Returns (cost, gapped, slot1, slot2, cells)."""
        a1, a2 = _as_u32(in1), _as_u32(in2)
        cap = len(a1) + len(a2) + 2
        g, s1, s2 = (np.zeros(cap, np.uint32) for _ in range(3))
        n = C.c_size_t(0)
        cells = C.c_uint64(0)
        cost = self.lib.oracle_align2d(self.t, _ptr(a1), len(a1), _ptr(a2), len(a2),
                                       _ptr(g), _ptr(s1), _ptr(s2), C.byref(n), C.byref(cells))
        k = n.value
        return int(cost), g[:k].copy(), s1[:k].copy(), s2[:k].copy(), int(cells.value)

    def align2d_affine(self, in1, in2, variant: int = 0):
        """Raw ``align2dAffine(..., getMedians=1)``; variant 0 = as coded (x86), 1 = patched shift.
        Returns (cost, gapped, slot1, slot2, cells)."""
        L = self.lib
        L.oracle_align2d_affine.restype = C.c_int64
        L.oracle_align2d_affine.argtypes = [C.c_void_p, C.c_int, _u32p, C.c_size_t, _u32p, C.c_size_t,
                                            _u32p, _u32p, _u32p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        a1, a2 = _as_u32(in1), _as_u32(in2)
        cap = len(a1) + len(a2) + 2
        g, s1, s2 = (np.zeros(cap, np.uint32) for _ in range(3))
        n = C.c_size_t(0)
        cells = C.c_uint64(0)
        cost = L.oracle_align2d_affine(self.t, variant, _ptr(a1), len(a1), _ptr(a2), len(a2),
                                       _ptr(g), _ptr(s1), _ptr(s2), C.byref(n), C.byref(cells))
        k = n.value
        return int(cost), g[:k].copy(), s1[:k].copy(), s2[:k].copy(), int(cells.value)

    def pairwise_do(self, c1, c2):
        """Haskell-level ``foreignPairwiseDO``. A character is ``None`` (Missing) or an (n,3) array of
        (median,left,right). Returns (cost, character-or-None, cells)."""
        def split(c):
            if c is None:
                z = np.zeros(1, np.uint32)
                return z, z, z, 0, 1
            c = _as_u32(c).reshape(-1, 3)
            return (np.ascontiguousarray(c[:, 0]), np.ascontiguousarray(c[:, 1]),
                    np.ascontiguousarray(c[:, 2]), len(c), 0)
        m1, l1, r1, n1, x1 = split(c1)
        m2, l2, r2, n2, x2 = split(c2)
        cap = n1 + n2 + 2
        om, ol, orr = (np.zeros(cap, np.uint32) for _ in range(3))
        on = C.c_size_t(0)
        omiss = C.c_int(0)
        cells = C.c_uint64(0)
        cost = self.lib.oracle_foreign_pairwise_do(
            self.t, _ptr(m1), _ptr(l1), _ptr(r1), n1, x1, _ptr(m2), _ptr(l2), _ptr(r2), n2, x2,
            _ptr(om), _ptr(ol), _ptr(orr), C.byref(on), C.byref(omiss), C.byref(cells))
        if omiss.value:
            return int(cost), None, int(cells.value)
        k = on.value
        return int(cost), np.stack([om[:k], ol[:k], orr[:k]], axis=1), int(cells.value)


def reroot(oracle: "Oracle", children, leaves, trial=None):
    """Re-rooting pass of one tree / one character (reroot_oracle.c).  children: (n_taxa-1, 2) int32, leaves: list of
    element strings.  Returns (min cost, edges [(u, v)], edge_cost, edge_local); with `trial` (a new leaf's string)
    a fifth item: the Wagner edge-trial cost per edge."""
    lib = oracle.lib
    lib.oracle_reroot_trial.restype = C.c_int64
    lib.oracle_reroot_trial.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 9 + [C.c_size_t, C.c_void_p]
    ch = np.ascontiguousarray(np.asarray(children, np.int32))
    n_taxa = ch.shape[0] + 1
    arrs = [_as_u32(x if x is not None else np.zeros(1, np.uint32)) for x in leaves]     # None = a Missing character
    ptrs = (C.c_void_p * n_taxa)(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * n_taxa)(*[len(a) if x is not None else C.c_size_t(-1).value for a, x in zip(arrs, leaves)])
    ne = 2 * n_taxa - 3
    eu, ev = np.zeros(ne, np.int32), np.zeros(ne, np.int32)
    ec, el = np.zeros(ne, np.int64), np.zeros(ne, np.int64)
    tr = _as_u32(trial) if trial is not None else None
    et = np.zeros(ne, np.int64)
    best = lib.oracle_reroot_trial(oracle.t, n_taxa, ch.ctypes.data, ptrs, lens, eu.ctypes.data, ev.ctypes.data,
                                   ec.ctypes.data, el.ctypes.data, None, tr.ctypes.data if tr is not None else None,
                                   len(tr) if tr is not None else 0, et.ctypes.data)
    if trial is None:
        return int(best), list(zip(eu.tolist(), ev.tolist())), ec, el
    return int(best), list(zip(eu.tolist(), ev.tolist())), ec, el, et


def preorder_tree(oracle: "Oracle", children, leaves):
    """Post-order contexts and pre-order implied alignments of one rooted tree / one character (preorder_oracle.c).
    Returns (root id, contexts, implied alignments, single disambiguations): dicts node -> (n, 3) uint32 array
    (median, left, right) / (n,) array, None for a Missing character."""
    lib = oracle.lib
    lib.oracle_preorder_tree.restype = C.c_int
    lib.oracle_preorder_tree.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_void_p] * 9
    ch = np.ascontiguousarray(np.asarray(children, np.int32))
    n_taxa = ch.shape[0] + 1
    n_nodes = 2 * n_taxa - 1
    arrs = [_as_u32(x) for x in leaves]
    ptrs = (C.c_void_p * n_taxa)(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * n_taxa)(*[len(a) for a in arrs])
    cap = sum(len(a) for a in arrs) + 2 * n_nodes + 8
    bufs = [np.zeros(n_nodes * cap, np.uint32) for _ in range(7)]
    ctx_len, ia_len = np.zeros(n_nodes, np.uint32), np.zeros(n_nodes, np.uint32)
    cm, cl, cr, im, il, ir, sg = bufs
    root = lib.oracle_preorder_tree(oracle.t, n_taxa, ch.ctypes.data, ptrs, lens, cap, cm.ctypes.data, cl.ctypes.data,
                                    cr.ctypes.data, ctx_len.ctypes.data, im.ctypes.data, il.ctypes.data, ir.ctypes.data,
                                    ia_len.ctypes.data, sg.ctypes.data)
    if root < 0:
        raise RuntimeError(f"oracle_preorder_tree failed: {root}")
    MISSING = 0xffffffff

    def triple(a, b, c, v, n):
        return None if n == MISSING else np.stack([a[v * cap:v * cap + n], b[v * cap:v * cap + n], c[v * cap:v * cap + n]], axis=1)
    ctxs = {v: triple(cm, cl, cr, v, int(ctx_len[v])) for v in range(n_nodes)}
    ias = {v: triple(im, il, ir, v, int(ia_len[v])) for v in range(n_nodes)}
    single = {v: None if int(ia_len[v]) == MISSING else sg[v * cap:v * cap + int(ia_len[v])].copy() for v in range(n_nodes)}
    return root, ctxs, ias, single


def reroot_haskell(dialect: int, a: int, children, leaves, sigma=None, metric_kind=None):
    """Re-rooting pass with every alignment by a restated Haskell dialect (alphabets > 8).
    metric_kind: None -> 0 (discrete) without sigma / 1 (explicit) with; 2 = L1 norm (firstLinearNormPairwiseLogic).
    Returns (min cost, edges, edge_cost, edge_local)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.oracle_reroot_haskell.restype = C.c_int64
    lib.oracle_reroot_haskell.argtypes = [C.c_int, C.c_int, _u32p, C.c_int, C.c_uint32] + [C.c_void_p] * 7
    ch = np.ascontiguousarray(np.asarray(children, np.int32))
    n_taxa = ch.shape[0] + 1
    W = (a + 31) // 32                      # leaves: (n, W) uint32 arrays when the alphabet exceeds 32 symbols
    arrs = [_as_u32(x if x is not None else np.zeros(W, np.uint32)).reshape(-1) for x in leaves]    # None = Missing
    ptrs = (C.c_void_p * n_taxa)(*[x.ctypes.data for x in arrs])
    lens = (C.c_size_t * n_taxa)(*[len(a) // W if x is not None else C.c_size_t(-1).value for a, x in zip(arrs, leaves)])
    ne = 2 * n_taxa - 3
    eu, ev = np.zeros(ne, np.int32), np.zeros(ne, np.int32)
    ec, el = np.zeros(ne, np.int64), np.zeros(ne, np.int64)
    sg = _as_u32(sigma) if sigma is not None else np.zeros(1, np.uint32)
    kind = metric_kind if metric_kind is not None else (0 if sigma is None else 1)
    best = lib.oracle_reroot_haskell(dialect, kind, _ptr(sg), a, n_taxa, ch.ctypes.data, ptrs, lens,
                                     eu.ctypes.data, ev.ctypes.data, ec.ctypes.data, el.ctypes.data)
    return int(best), list(zip(eu.tolist(), ev.tolist())), ec, el


def haskell_do(dialect: int, a: int, in1, in2, sigma=None, metric_kind=None):
    """S1 / S2 Haskell dialects on median strings (see do_oracle.c). sigma=None -> discrete metric; metric_kind 2 = the L1
    closed form.  Returns (cost, median, aligned1, aligned2, cells, final_offset)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.oracle_haskell_do.restype = C.c_int64
    lib.oracle_haskell_do.argtypes = [C.c_int, C.c_int, _u32p, C.c_int, _u32p, C.c_size_t, _u32p, C.c_size_t,
                                      _u32p, _u32p, _u32p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_int)]
    a1, a2 = _as_u32(in1), _as_u32(in2)
    sg = _as_u32(sigma) if sigma is not None else np.zeros(1, np.uint32)
    cap = len(a1) + len(a2) + 2
    m, x, y = (np.zeros(cap, np.uint32) for _ in range(3))
    n = C.c_size_t(0)
    cells = C.c_uint64(0)
    off = C.c_int(0)
    kind = metric_kind if metric_kind is not None else (0 if sigma is None else 1)
    cost = lib.oracle_haskell_do(dialect, kind, _ptr(sg), a, _ptr(a1), len(a1), _ptr(a2), len(a2),
                                 _ptr(m), _ptr(x), _ptr(y), C.byref(n), C.byref(cells), C.byref(off))
    k = n.value
    return int(cost), m[:k].copy(), x[:k].copy(), y[:k].copy(), int(cells.value), int(off.value)


def overlap(a: int, x: int, y: int, sigma=None):
    """(median, cost) of one overlap lookup by the restatement (sigma=None -> discrete metric)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.oracle_overlap.restype = C.c_uint64
    lib.oracle_overlap.argtypes = [C.c_int, _u32p, C.c_int, C.c_uint32, C.c_uint32, _u32p]
    sg = _as_u32(sigma) if sigma is not None else np.zeros(1, np.uint32)
    med = C.c_uint32(0)
    cost = lib.oracle_overlap(0 if sigma is None else 1, _ptr(sg), a, x, y, C.byref(med))
    return int(med.value), int(cost)


class _DcElement(C.Structure):
    _fields_ = [("alphSize", C.c_size_t), ("element", C.POINTER(C.c_uint64))]


class ReferenceMemo:
    """The compiled, unmodified reference tcm-memo (``oracle/_ref/libtcmmemo_ref.so``), driven like
    ``getMedianAndCost2D`` (lib/tcm-memo/src/Data/TCM/Memoized/FFI.hsc:97-112): malloc'ed dcElement_t buffers."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_MEMO_SO)

    def __init__(self, sigma):
        self.lib = C.CDLL(REF_MEMO_SO)
        self.libc = C.CDLL(None)
        self.libc.calloc.restype = C.c_void_p
        self.libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
        self.libc.free.argtypes = [C.c_void_p]
        self.sigma = _as_u32(sigma).copy()
        self.a = int(self.sigma.shape[0])
        self.lib.matrixInit.restype = C.c_void_p
        self.lib.matrixInit.argtypes = [C.c_size_t, _u32p]
        self.lib.getCostAndMedian2D.restype = C.c_uint
        self.lib.getCostAndMedian2D.argtypes = [C.POINTER(_DcElement)] * 3 + [C.c_void_p]
        self.m = self.lib.matrixInit(self.a, _ptr(self.sigma))

    def _elem(self, v):
        buf = self.libc.calloc(1, 8)
        C.cast(buf, C.POINTER(C.c_uint64))[0] = int(v)
        return _DcElement(self.a, C.cast(buf, C.POINTER(C.c_uint64)))

    def overlap(self, x: int, y: int):
        e1, e2, r = self._elem(x), self._elem(y), self._elem(0)
        cost = self.lib.getCostAndMedian2D(C.byref(e1), C.byref(e2), C.byref(r), self.m)
        med = int(r.element[0])
        for e in (e1, e2, r):
            self.libc.free(C.cast(e.element, C.c_void_p))
        return med, int(cost)


class _AlignIO(C.Structure):
    _fields_ = [("character", _u32p), ("length", C.c_size_t), ("capacity", C.c_size_t)]


class Reference:
    """The compiled, unmodified reference aligner (``oracle/_ref/libpcgdo_ref.so``)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self, sigma, gap_open: int = 0, patched: bool = False):
        self.lib = C.CDLL(REF_PATCHED_SO if patched else REF_SO)
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        self.libc.free.argtypes = [C.c_void_p]
        sigma = _as_u32(sigma)
        self.a = int(sigma.shape[0])
        self.gap = 1 << (self.a - 1)
        self.cm = C.create_string_buffer(256)       # cost_matrices_2d_t (EXT/costMatrix.h:35-93) fits
        self.lib.setUp2dCostMtx.argtypes = [C.c_void_p, _u32p, C.c_size_t, C.c_uint]
        self.lib.setUp2dCostMtx(self.cm, _ptr(sigma), self.a, gap_open)
        self.lib.align2d.restype = C.c_int
        self.lib.align2d.argtypes = [C.POINTER(_AlignIO)] * 4 + [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.align2dAffine.restype = C.c_int
        self.lib.align2dAffine.argtypes = [C.POINTER(_AlignIO)] * 4 + [C.c_void_p, C.c_int]

    def tables(self):
        """(cost, median) dense tables as the reference built them."""
        class _CM(C.Structure):
            _fields_ = [("alphSize", C.c_size_t), ("costMatrixDimension", C.c_size_t), ("gap_char", C.c_uint),
                        ("cost_model_type", C.c_int), ("include_ambiguities", C.c_int),
                        ("gap_open_cost", C.c_uint), ("is_metric", C.c_int), ("num_elements", C.c_size_t),
                        ("cost", _u32p), ("median", _u32p), ("worst", _u32p),
                        ("prepend_cost", _u32p), ("tail_cost", _u32p)]
        cm = C.cast(self.cm, C.POINTER(_CM)).contents
        dim = 1 << self.a
        return (np.ctypeslib.as_array(cm.cost, shape=(dim, dim)).copy(),
                np.ctypeslib.as_array(cm.median, shape=(dim, dim)).copy())

    def _aio(self, vals, cap):
        """allocInitAlign_io (FFI.hsc:474-481): malloc'ed buffer, payload in the LAST len slots."""
        n = len(vals)
        buf = self.libc.malloc(cap * 4)
        arr = (C.c_uint32 * cap).from_address(buf)
        C.memset(buf, 0, cap * 4)
        for i, v in enumerate(vals):
            arr[cap - n + i] = int(v)
        return _AlignIO(C.cast(buf, _u32p), n, cap)

    def _take(self, aio):
        off = aio.capacity - aio.length
        out = np.array([aio.character[off + i] for i in range(aio.length)], dtype=np.uint32)
        self.libc.free(C.cast(aio.character, C.c_void_p))
        return out

    def align2d(self, in1, in2, affine: bool = False):
        cap = len(in1) + len(in2) + 2
        a1, a2 = self._aio(in1, cap), self._aio(in2, cap)
        g, u = self._aio([], cap), self._aio([], cap)
        if affine:
            cost = self.lib.align2dAffine(C.byref(a1), C.byref(a2), C.byref(g), C.byref(u), self.cm, 1)
        else:
            cost = self.lib.align2d(C.byref(a1), C.byref(a2), C.byref(g), C.byref(u), self.cm, 0, 1, 0)
        s1, s2, gg = self._take(a1), self._take(a2), self._take(g)
        self.libc.free(C.cast(u.character, C.c_void_p))
        return int(cost), gg, s1, s2


def cpu_bench_align2d(sigma, pp, threads: int, kind: str = "reference", want_cells: bool = True,
                      gap_open: int = 0, affine_variant: int = 0):
    """Time the CPU path on `threads` host threads over the packed pairs `pp` (a pcg_b200 PackedPairs-like
    object: elems/off1/len1/off2/len2/n).  kind: "reference" (oracle/_ref) or "port" (do_oracle.c).
    Returns (seconds, cost[int32], cells[uint64])."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.cpu_bench_align2d_ex.restype = C.c_int
    lib.cpu_bench_align2d_ex.argtypes = [C.c_char_p, _u32p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                         C.c_void_p, C.POINTER(C.c_double)]
    sigma = _as_u32(sigma)
    cost = np.zeros(pp.n, np.int32)
    cells = np.zeros(pp.n, np.uint64)
    sec = C.c_double(0)
    ref_path = REF_PATCHED_SO if (gap_open and affine_variant == 1) else REF_SO
    so = ref_path.encode() if kind == "reference" else None
    if kind == "reference" and not os.path.exists(ref_path):
        raise FileNotFoundError(ref_path)
    rc = lib.cpu_bench_align2d_ex(so, _ptr(sigma), int(sigma.shape[0]), gap_open, affine_variant, pp.elems.ctypes.data,
                                  pp.off1.ctypes.data, pp.len1.ctypes.data, pp.off2.ctypes.data, pp.len2.ctypes.data,
                                  pp.n, threads, cost.ctypes.data, cells.ctypes.data if want_cells else None,
                                  C.byref(sec))
    if rc:
        raise RuntimeError(f"cpu_bench_align2d failed: {rc}")
    return float(sec.value), cost, cells


def cpu_bench_postorder(sigma, children, n_taxa, n_chars, leaf_elems, leaf_off, leaf_len, col0, n_cols,
                        threads: int, kind: str = "reference"):
    """CPU post-order over columns col0..col0+n_cols-1 of the (tree x character) grid.
    Returns (seconds, col_cost[int64], col_cells[uint64])."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.cpu_bench_postorder.restype = C.c_int
    lib.cpu_bench_postorder.argtypes = [C.c_char_p, _u32p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    sigma = _as_u32(sigma)
    ch = np.ascontiguousarray(np.asarray(children, np.int32))
    cost = np.zeros(n_cols, np.int64)
    cells = np.zeros(n_cols, np.uint64)
    sec = C.c_double(0)
    so = REF_SO.encode() if kind == "reference" else None
    if kind == "reference" and not os.path.exists(REF_SO):
        raise FileNotFoundError(REF_SO)
    rc = lib.cpu_bench_postorder(so, _ptr(sigma), int(sigma.shape[0]), ch.shape[0], n_taxa, n_chars, ch.ctypes.data,
                                 leaf_elems.ctypes.data, leaf_off.ctypes.data, leaf_len.ctypes.data, col0, n_cols,
                                 threads, cost.ctypes.data, cells.ctypes.data, C.byref(sec))
    if rc:
        raise RuntimeError(f"cpu_bench_postorder failed: {rc}")
    return float(sec.value), cost, cells


def cpu_bench_haskell(dialect: int, a: int, elems, off1, len1, off2, len2, threads: int, sigma=None,
                      through_reference_memo: bool = False):
    """Time the restated Haskell dialect on `threads` host threads (elems: uint32 ambiguity sets).
    sigma=None -> discrete metric; through_reference_memo -> every overlap answered by the compiled reference
    tcm-memo.  Returns (seconds, cost[int32], cells[uint64], memo lookups)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.cpu_bench_haskell.restype = C.c_int
    lib.cpu_bench_haskell.argtypes = [C.c_char_p, C.c_int, C.c_int, _u32p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    n = len(len1)
    sg = _as_u32(sigma) if sigma is not None else np.zeros(1, np.uint32)
    cost = np.zeros(n, np.int32)
    cells = np.zeros(n, np.uint64)
    sec = C.c_double(0)
    lk = C.c_uint64(0)
    so = None
    if through_reference_memo:
        if not os.path.exists(REF_MEMO_SO):
            raise FileNotFoundError(REF_MEMO_SO)
        so = REF_MEMO_SO.encode()
    rc = lib.cpu_bench_haskell(so, dialect, 0 if sigma is None else 1, _ptr(sg), a, elems.ctypes.data, off1.ctypes.data,
                               len1.ctypes.data, off2.ctypes.data, len2.ctypes.data, n, threads, cost.ctypes.data,
                               cells.ctypes.data, C.byref(sec), C.byref(lk))
    if rc:
        raise RuntimeError(f"cpu_bench_haskell failed: {rc}")
    return float(sec.value), cost, cells, int(lk.value)


def haskell_do_w(dialect: int, a: int, in1, in2, sigma=None, metric_kind=None):
    """As haskell_do for alphabets of more than 32 symbols: in1 / in2 are (n, W) uint32 arrays, W = ceil(a / 32).
    Returns (cost, median (k, W), aligned1, aligned2, cells, final_offset)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.oracle_haskell_do_w.restype = C.c_int64
    lib.oracle_haskell_do_w.argtypes = [C.c_int, C.c_int, _u32p, C.c_int, _u32p, C.c_size_t, _u32p, C.c_size_t,
                                        _u32p, _u32p, _u32p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_int)]
    W = (a + 31) // 32
    a1, a2 = _as_u32(in1).reshape(-1, W), _as_u32(in2).reshape(-1, W)
    sg = _as_u32(sigma) if sigma is not None else np.zeros(1, np.uint32)
    cap = len(a1) + len(a2) + 2
    m, x, y = (np.zeros((cap, W), np.uint32) for _ in range(3))
    n = C.c_size_t(0)
    cells = C.c_uint64(0)
    off = C.c_int(0)
    kind = metric_kind if metric_kind is not None else (0 if sigma is None else 1)
    cost = lib.oracle_haskell_do_w(dialect, kind, _ptr(sg), a, _ptr(a1), len(a1), _ptr(a2), len(a2),
                                   _ptr(m), _ptr(x), _ptr(y), C.byref(n), C.byref(cells), C.byref(off))
    k = n.value
    return int(cost), m[:k].copy(), x[:k].copy(), y[:k].copy(), int(cells.value), int(off.value)


# ------------------------------------------------------------------------------------------------
# Whole-batch parity helpers: the CPU path over many pairs on host threads, returning per pair the alignment
# length and the zlib-compatible CRC-32 of each of the three output strings (cpu_bench.c), so that tests can
# compare thousands of complete alignments with the compiled reference without shipping the strings.
# ------------------------------------------------------------------------------------------------
def cpu_align2d_full(sigma, pp, threads: int, kind: str = "reference", gap_open: int = 0, affine_variant: int = 0):
    """Returns (cost[int32], out_len[uint32], crc[uint32 (n, 3)] of gapped / slot1 / slot2 as one byte per element)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.cpu_bench_align2d_full.restype = C.c_int
    lib.cpu_bench_align2d_full.argtypes = [C.c_char_p, _u32p, C.c_int, C.c_uint32, C.c_int] + [C.c_void_p] * 5 + \
                                          [C.c_size_t, C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_double)]
    sigma = _as_u32(sigma)
    cost = np.zeros(pp.n, np.int32)
    out_len = np.zeros(pp.n, np.uint32)
    crc = np.zeros((pp.n, 3), np.uint32)
    sec = C.c_double(0)
    ref_path = REF_PATCHED_SO if (gap_open and affine_variant == 1) else REF_SO
    so = ref_path.encode() if kind == "reference" else None
    if kind == "reference" and not os.path.exists(ref_path):
        raise FileNotFoundError(ref_path)
    rc = lib.cpu_bench_align2d_full(so, _ptr(sigma), int(sigma.shape[0]), gap_open, affine_variant, pp.elems.ctypes.data,
                                    pp.off1.ctypes.data, pp.len1.ctypes.data, pp.off2.ctypes.data, pp.len2.ctypes.data,
                                    pp.n, threads, cost.ctypes.data, None, out_len.ctypes.data, crc.ctypes.data,
                                    C.byref(sec))
    if rc:
        raise RuntimeError(f"cpu_bench_align2d_full failed: {rc}")
    return cost, out_len, crc


def cpu_postorder_full(sigma, children, n_taxa, n_chars, leaf_elems, leaf_off, leaf_len, col0, n_cols, threads: int,
                       kind: str = "reference"):
    """Post-order over columns col0.. of the (tree x character) grid; returns (col_cost[int64], node_total[int64
    (n_cols, n_taxa-1)], node_local[int32 same], node_crc[uint32 same]: CRC-32 of each node's ungapped median)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.cpu_bench_postorder_full.restype = C.c_int
    lib.cpu_bench_postorder_full.argtypes = [C.c_char_p, _u32p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32] + \
                                            [C.c_void_p] * 4 + [C.c_uint32, C.c_uint32, C.c_int] + [C.c_void_p] * 5 + \
                                            [C.POINTER(C.c_double)]
    sigma = _as_u32(sigma)
    ch = np.ascontiguousarray(np.asarray(children, np.int32))
    ni = n_taxa - 1
    cost = np.zeros(n_cols, np.int64)
    tot = np.zeros((n_cols, ni), np.int64)
    loc = np.zeros((n_cols, ni), np.int32)
    crc = np.zeros((n_cols, ni), np.uint32)
    sec = C.c_double(0)
    so = REF_SO.encode() if kind == "reference" else None
    if kind == "reference" and not os.path.exists(REF_SO):
        raise FileNotFoundError(REF_SO)
    rc = lib.cpu_bench_postorder_full(so, _ptr(sigma), int(sigma.shape[0]), ch.shape[0], n_taxa, n_chars, ch.ctypes.data,
                                      leaf_elems.ctypes.data, leaf_off.ctypes.data, leaf_len.ctypes.data, col0, n_cols,
                                      threads, cost.ctypes.data, None, tot.ctypes.data, loc.ctypes.data, crc.ctypes.data,
                                      C.byref(sec))
    if rc:
        raise RuntimeError(f"cpu_bench_postorder_full failed: {rc}")
    return cost, tot, loc, crc


def cpu_haskell_full(dialect: int, a: int, elems, off1, len1, off2, len2, threads: int, sigma=None,
                     through_reference_memo: bool = False):
    """The restated Haskell dialect over many pairs (a <= 32); every overlap answered by the compiled reference
    tcm-memo when `through_reference_memo`.  Returns (cost, cells, out_len, crc (n, 3) over uint32 words, final_offset)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    lib = C.CDLL(ORACLE_SO)
    lib.cpu_bench_haskell_full.restype = C.c_int
    lib.cpu_bench_haskell_full.argtypes = [C.c_char_p, C.c_int, C.c_int, _u32p, C.c_int] + [C.c_void_p] * 5 + \
                                          [C.c_size_t, C.c_int] + [C.c_void_p] * 5 + \
                                          [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    n = len(len1)
    sg = _as_u32(sigma) if sigma is not None else np.zeros(1, np.uint32)
    cost = np.zeros(n, np.int32)
    cells = np.zeros(n, np.uint64)
    out_len = np.zeros(n, np.uint32)
    crc = np.zeros((n, 3), np.uint32)
    fo = np.zeros(n, np.int32)
    sec = C.c_double(0)
    lk = C.c_uint64(0)
    so = None
    if through_reference_memo:
        if not os.path.exists(REF_MEMO_SO):
            raise FileNotFoundError(REF_MEMO_SO)
        so = REF_MEMO_SO.encode()
    rc = lib.cpu_bench_haskell_full(so, dialect, 0 if sigma is None else 1, _ptr(sg), a, elems.ctypes.data, off1.ctypes.data,
                                    len1.ctypes.data, off2.ctypes.data, len2.ctypes.data, n, threads, cost.ctypes.data,
                                    cells.ctypes.data, out_len.ctypes.data, crc.ctypes.data, fo.ctypes.data,
                                    C.byref(sec), C.byref(lk))
    if rc:
        raise RuntimeError(f"cpu_bench_haskell_full failed: {rc}")
    return cost, cells, out_len, crc, fo
