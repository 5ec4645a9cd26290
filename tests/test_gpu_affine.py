"""GPU parity: C-affine dialect (align2dAffine) against the oracle restatement, which is itself checked against
the compiled reference (as coded) and a one-expression-patched build of it (tests/test_oracle.py).
PARITY UNPINNED upstream: the reference has no golden vector for affine costs."""
import numpy as np
import pytest

from tests.helpers import random_pair, random_string, tcm_matrix

pytestmark = pytest.mark.gpu


def _check(o, ctx, tcm, pairs, variant):
    import pcg_b200 as P
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    r = P.align2d_batch(ctx, tcm, pp, want_cells=True)
    for k, (s1, s2) in enumerate(pairs):
        oc = o.align2d_affine(s1, s2, variant)
        gc = r.pair(pp, k)
        assert gc[0] == oc[0], (k, len(s1), len(s2), gc[0], oc[0])
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), (k, len(s1), len(s2))
        assert int(r.cells[k]) == oc[4], (k, len(s1), len(s2))


@pytest.mark.parametrize("name,go,variant", [("discrete", 3, 0), ("l1", 3, 0), ("discrete", 1, 1), ("l1", 3, 1),
                                             ("weird", 2, 1)])
def test_affine_fuzz(oracle_mod, gpu_ctx, name, go, variant):
    sig = tcm_matrix(name)
    o = oracle_mod.Oracle(sig, go)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=700, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig, gap_open=go, affine_variant=variant)
    rng = np.random.default_rng(300 + go)
    pairs = [random_pair(rng, 300, amb=0.2) for _ in range(150)]
    for n1, n2 in [(1, 1), (2, 1), (1, 30), (39, 39), (40, 40), (41, 41), (42, 40), (80, 41), (200, 35), (600, 90),
                   (130, 131), (640, 600)]:
        pairs.append((random_string(rng, n1, amb=0.2), random_string(rng, n2, amb=0.2)))
    s = random_string(rng, 90, amb=0.3)
    pairs.append((s, s.copy()))
    _check(o, gpu_ctx, tcm, pairs, variant)


def test_affine_oracle_matches_reference_builds(oracle_mod):
    """(CPU part, kept beside the GPU test so the chain of evidence reads in one place.)"""
    if not oracle_mod.Reference.available():
        pytest.skip("oracle/_ref not built")
    import os
    for variant, patched in [(0, False), (1, True)]:
        if patched and not os.path.exists(oracle_mod.REF_PATCHED_SO):
            continue
        sig = tcm_matrix("l1")
        o, r = oracle_mod.Oracle(sig, 3), oracle_mod.Reference(sig, 3, patched=patched)
        rng = np.random.default_rng(8)
        for _ in range(60):
            s1, s2 = random_pair(rng, 200, amb=0.2)
            rc, oc = r.align2d(s1, s2, affine=True), o.align2d_affine(s1, s2, variant)
            assert rc[0] == oc[0]
            for x, y in zip(rc[1:], oc[1:4]):
                assert np.array_equal(x, y)


def test_reference_compatible_align2dAffine_symbol(oracle_mod, gpu_ctx):
    """`align2dAffine` called the way FFI.hsc does, against the compiled reference's own answer (as coded)."""
    import ctypes as C
    import pcg_b200.backend as be
    lib = be.load()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    sig = tcm_matrix("l1")
    o = oracle_mod.Oracle(sig, 3)
    cm = be.CostMatrices2D()
    lib.setUp2dCostMtx(C.byref(cm), sig.ctypes.data_as(C.POINTER(C.c_uint32)), 5, 3)
    assert cm.cost_model_type == 3 and cm.gap_open_cost == 3

    def aio(vals, cap):
        buf = libc.malloc(cap * 4)
        C.memset(buf, 0, cap * 4)
        arr = (C.c_uint32 * cap).from_address(buf)
        for i, v in enumerate(vals):
            arr[cap - len(vals) + i] = int(v)
        return be.AlignIO(C.cast(buf, C.POINTER(C.c_uint32)), len(vals), cap)

    def take(io):
        off = io.capacity - io.length
        out = [io.character[off + i] for i in range(io.length)]
        libc.free(C.cast(io.character, C.c_void_p))
        return out

    rng = np.random.default_rng(77)
    for _ in range(5):
        s1, s2 = random_pair(rng, 150, amb=0.2)
        cap = len(s1) + len(s2) + 2
        a1, a2, g, u = aio(s1, cap), aio(s2, cap), aio([], cap), aio([], cap)
        cost = lib.align2dAffine(C.byref(a1), C.byref(a2), C.byref(g), C.byref(u), C.byref(cm), 1)
        oc = o.align2d_affine(s1, s2, 0)
        assert cost == oc[0]
        assert take(g) == oc[1].tolist() and take(a1) == oc[2].tolist() and take(a2) == oc[3].tolist()
        libc.free(C.cast(u.character, C.c_void_p))
