"""The drop-in boundary beyond the 2D aligner: the 3D setup / lookup symbols PCG's Haskell side links
(Data/TCM/Dense/FFI.hsc:327-334, Data/TCM/Memoized/FFI.hsc:63-69), wide alphabets through the tcm-memo symbols, and the
link proof of INTEGRATION.md §1 (a C stand-in for the seven foreign imports linked against this library alone)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.helpers import ROOT, tcm_matrix


@pytest.mark.parametrize("name,a,go", [("discrete", 5, 0), ("l1", 5, 3), ("weird", 5, 0), ("1:2", 4, 0), ("l1", 6, 0)])
def test_setUp3dCostMtx_equals_compiled_reference(oracle_mod, name, a, go):
    """Every entry of both three-way tables against setUp3dCostMtx of the unmodified reference (oracle/_ref)."""
    if not oracle_mod.Reference.available():
        pytest.skip("oracle/_ref not built")
    import pcg_b200.backend as be
    lib = be.load()
    ref = C.CDLL(oracle_mod.REF_SO)
    sig = tcm_matrix(name, a)
    mine, theirs = be.CostMatrices3D(), be.CostMatrices3D()
    lib.setUp3dCostMtx(C.byref(mine), sig.ctypes.data_as(C.POINTER(C.c_uint32)), a, go)
    ref.setUp3dCostMtx.argtypes = [C.POINTER(be.CostMatrices3D), C.POINTER(C.c_uint32), C.c_size_t, C.c_uint]
    ref.setUp3dCostMtx.restype = None
    ref.setUp3dCostMtx(C.byref(theirs), sig.ctypes.data_as(C.POINTER(C.c_uint32)), a, go)
    for f in ("alphSize", "costMatrixDimension", "gap_char", "cost_model_type", "include_ambiguities", "gap_open_cost",
              "num_elements"):
        assert getattr(mine, f) == getattr(theirs, f), f
    n = ((1 << a) + 1) ** 3
    assert np.array_equal(np.ctypeslib.as_array(mine.cost, shape=(n,)), np.ctypeslib.as_array(theirs.cost, shape=(n,)))
    assert np.array_equal(np.ctypeslib.as_array(mine.median, shape=(n,)), np.ctypeslib.as_array(theirs.median, shape=(n,)))
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for m in (mine, theirs):
        libc.free(C.cast(m.cost, C.c_void_p)); libc.free(C.cast(m.median, C.c_void_p))


def test_struct_layout_3d():
    import pcg_b200.backend as be
    # cost_matrices_3d_t (costMatrix.h:102-130): 9 fields; hsc pokes `cost` / `median` at these offsets (Dense/FFI.hsc:236-300)
    assert C.sizeof(be.CostMatrices3D) == 56
    assert be.CostMatrices3D.cost.offset == 40 and be.CostMatrices3D.median.offset == 48


def _link(tmp_path):
    import pcg_b200.backend as be
    be.load()
    exe = str(tmp_path / "hs_imports")
    libdir = os.path.join(ROOT, "pcg_b200", "lib")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "link", "hs_imports.c"), "-o", exe, "-L", libdir, "-lpcgdo_b200",
           "-Wl,--no-undefined", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True)
    return exe


def test_drop_in_recipe_links_without_reference_objects(tmp_path, oracle_mod):
    """all eight EXT c-sources and the four tcm-memo sources removed, libpcgdo_b200.so linked in
    their place.  The stand-in for the Haskell imports links with --no-undefined and its host-only calls reproduce
    the reference's tables (same checksum program run against oracle/_ref when that exists)."""
    exe = _link(tmp_path)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == "tables"
    if oracle_mod.Reference.available():
        import pcg_b200.backend as be
        ref = C.CDLL(oracle_mod.REF_SO)
        a, dim = 5, 32
        sig = tcm_matrix("discrete", a)
        m2, m3 = be.CostMatrices2D(), be.CostMatrices3D()
        ref.setUp2dCostMtx(C.byref(m2), sig.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(a), C.c_uint(0))
        ref.setUp3dCostMtx(C.byref(m3), sig.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(a), C.c_uint(0))

        def chk(cost, median, n):
            s = 0
            for c, m in zip(np.ctypeslib.as_array(cost, shape=(n,)).tolist(), np.ctypeslib.as_array(median, shape=(n,)).tolist()):
                s = (s * 31 + c + 7 * m) & 0xffffffffffffffff
            return s
        assert int(out[1]) == chk(m2.cost, m2.median, dim * dim)
        assert int(out[2]) == chk(m3.cost, m3.median, (dim + 1) ** 3)


@pytest.mark.gpu
def test_drop_in_imports_run_on_the_gpu(tmp_path, gpu_ctx):
    exe = _link(tmp_path)
    out = subprocess.run([exe, "gpu"], check=True, capture_output=True, text=True).stdout.splitlines()
    assert out[1].startswith("align2d 2 len 7")          # one substitution + one indel under the discrete metric
    # sigma = |i - j| over 3 symbols: {0} vs {2} -> every symbol costs 2, median {0,1,2}; with {1} added: symbol 1 costs 2
    assert out[2] == "overlap2 2 7"
    assert out[3] == "overlap3 2 2"


@pytest.mark.gpu
@pytest.mark.parametrize("a", [21, 40, 82, 256])
def test_tcm_memo_symbols_on_wide_alphabets(oracle_mod, gpu_ctx, a):
    """matrixInit / getCostAndMedian2D / getCostAndMedian3D exactly as Memoized/FFI.hsc calls them (malloc'ed
    dcElement_t, callee re-allocates the result) for alphabets beyond 32 symbols, against the compiled reference
    tcm-memo answering the same calls."""
    if not oracle_mod.ReferenceMemo.available():
        pytest.skip("oracle/_ref not built")
    import pcg_b200.backend as be
    lib = be.load()
    ref = C.CDLL(oracle_mod.REF_MEMO_SO)
    libc = C.CDLL(None)
    libc.calloc.restype = C.c_void_p
    libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(a)
    sig = rng.integers(0, 9, size=(a, a)).astype(np.uint32)
    sig = ((sig + sig.T) // 2).astype(np.uint32)
    np.fill_diagonal(sig, 0)
    sig = np.ascontiguousarray(sig)
    for L in (lib, ref):
        L.matrixInit.restype = C.c_void_p
        L.matrixInit.argtypes = [C.c_size_t, C.POINTER(C.c_uint32)]
        L.getCostAndMedian2D.restype = C.c_uint
        L.getCostAndMedian2D.argtypes = [C.POINTER(be.DcElement)] * 3 + [C.c_void_p]
        L.getCostAndMedian3D.restype = C.c_uint
        L.getCostAndMedian3D.argtypes = [C.POINTER(be.DcElement)] * 4 + [C.c_void_p]
    mine, theirs = lib.matrixInit(a, sig.ctypes.data_as(C.POINTER(C.c_uint32))), ref.matrixInit(a, sig.ctypes.data_as(C.POINTER(C.c_uint32)))
    W64 = (a + 63) // 64

    def elem(bits):
        buf = libc.calloc(W64, 8)
        arr = C.cast(buf, C.POINTER(C.c_uint64))
        for b in bits:
            arr[b // 64] |= 1 << (b % 64)
        return be.DcElement(a, arr)

    def words(e):
        return [int(e.element[w]) for w in range(W64)]

    for t in range(40):
        sets = [sorted(set(rng.integers(0, a, size=int(rng.integers(1, 5))).tolist())) for _ in range(3)]
        got = []
        for L, m in ((lib, mine), (ref, theirs)):
            e = [elem(s) for s in sets]
            r2, r3 = elem([]), elem([])
            c2 = L.getCostAndMedian2D(C.byref(e[0]), C.byref(e[1]), C.byref(r2), m)
            c3 = L.getCostAndMedian3D(C.byref(e[0]), C.byref(e[1]), C.byref(e[2]), C.byref(r3), m)
            got.append((c2, words(r2), c3, words(r3)))
            for x in e + [r2, r3]:
                libc.free(C.cast(x.element, C.c_void_p))
        assert got[0] == got[1], (t, sets, got)
    lib.matrixDestroy(mine)


@pytest.mark.gpu
def test_shim_cache_survives_a_reused_address(oracle_mod, gpu_ctx):
    """freeCostMtx drops the cached device TCM, and a cost matrix rebuilt at the same address with other costs is
    re-uploaded (the cache verifies the symbol costs on every hit)."""
    import pcg_b200.backend as be
    lib = be.load()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(5)
    s1 = (1 << rng.integers(0, 4, size=60)).astype(np.uint32)
    s2 = (1 << rng.integers(0, 4, size=55)).astype(np.uint32)

    def aio(vals, cap):
        buf = libc.malloc(cap * 4)
        C.memset(buf, 0, cap * 4)
        arr = (C.c_uint32 * cap).from_address(buf)
        for i, v in enumerate(vals):
            arr[cap - len(vals) + i] = int(v)
        return be.AlignIO(C.cast(buf, C.POINTER(C.c_uint32)), len(vals), cap)

    def run(cm_ptr):
        cap = len(s1) + len(s2) + 2
        a1, a2, g, u = aio(s1, cap), aio(s2, cap), aio([], cap), aio([], cap)
        cost = lib.align2d(C.byref(a1), C.byref(a2), C.byref(g), C.byref(u), cm_ptr, 0, 1, 0)
        for x in (a1, a2, g, u):
            libc.free(C.cast(x.character, C.c_void_p))
        return cost

    seen = []
    for name in ("discrete", "l1", "discrete"):
        sig = tcm_matrix(name)
        cm = C.cast(libc.malloc(C.sizeof(be.CostMatrices2D)), C.POINTER(be.CostMatrices2D))
        lib.setUp2dCostMtx(cm, sig.ctypes.data_as(C.POINTER(C.c_uint32)), 5, 0)
        got = run(cm)
        assert got == oracle_mod.Oracle(sig).align2d(s1, s2)[0], name
        # overwrite the tables IN PLACE with another matrix's: same address, other costs
        other = tcm_matrix("l1" if name == "discrete" else "discrete")
        tmp = be.CostMatrices2D()
        lib.setUp2dCostMtx(C.byref(tmp), other.ctypes.data_as(C.POINTER(C.c_uint32)), 5, 0)
        C.memmove(cm.contents.cost, tmp.cost, 2 * 32 * 32 * 4)
        C.memmove(cm.contents.median, tmp.median, 2 * 32 * 32 * 4)
        assert run(cm) == oracle_mod.Oracle(other).align2d(s1, s2)[0], name
        seen.append(got)
        for f in ("cost", "median", "worst", "prepend_cost", "tail_cost"):
            libc.free(C.cast(getattr(tmp, f), C.c_void_p))
        lib.freeCostMtx(cm, 1)
    assert seen[0] == seen[2] != seen[1]
