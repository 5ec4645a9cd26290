"""GPU parity: explicit symbol-change matrices for alphabets the dense table cannot hold — the device-side
replacement of the reference's memoized TCM (lib/tcm-memo) and the Haskell dialects running over it.
The overlap is checked against committed answers of the compiled reference tcm-memo; the dialects against the
oracle's restatement (PARITY UNPINNED upstream for the dialect logic)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests.test_gpu_metric import _pairs, _rand_str

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tcm_memo_golden.json")


def _sigma(name, a, rng):
    if name == "pref-gap":                 # bench/strings/protein/protein-pref-gap.tcm shape: sub 2, indel 1
        m = 2 * (1 - np.eye(a, dtype=np.uint32))
        m[-1, :-1] = 1
        m[:-1, -1] = 1
        return m.astype(np.uint32)
    if name == "pref-sub":                 # sub 1, indel 2
        m = (1 - np.eye(a, dtype=np.uint32))
        m[-1, :-1] = 2
        m[:-1, -1] = 2
        return m.astype(np.uint32)
    if name == "random-sym":
        m = rng.integers(1, 9, size=(a, a))
        m = np.minimum(m, m.T)
        np.fill_diagonal(m, 0)
        return m.astype(np.uint32)
    if name == "random-asym":
        m = rng.integers(0, 7, size=(a, a))
        return m.astype(np.uint32)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["bench/strings/protein/protein-pref-gap.tcm", "bench/strings/protein/protein-pref-sub.tcm",
                                  "bench/strings/protein/protein-L1-norm.tcm", "bench/strings/dna/dna-pref-gap.tcm",
                                  "random-asymmetric-12"])
def test_device_overlap_reproduces_tcm_memo_golden(gpu_ctx, name):
    g = json.load(open(GOLD))[name]
    sg = np.array(g["sigma"], np.uint32)
    memo = gpu_ctx.memo(sg)
    rec = np.array(g["lookups"], np.int64)
    med, cost = memo.overlap(rec[:, 0], rec[:, 1])
    assert np.array_equal(med.astype(np.int64), rec[:, 2])
    assert np.array_equal(cost.astype(np.int64), rec[:, 3])


def test_reference_compatible_tcm_memo_symbols(gpu_ctx):
    """matrixInit / getCostAndMedian2D / matrixDestroy driven like lib/tcm-memo/src/Data/TCM/Memoized/FFI.hsc:97-112."""
    from pcg_b200.backend import DcElement, load
    lib = load()
    g = json.load(open(GOLD))["bench/strings/protein/protein-pref-gap.tcm"]
    sg = np.ascontiguousarray(np.array(g["sigma"], np.uint32))
    libc = C.CDLL(None)
    libc.calloc.restype = C.c_void_p
    libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    m = lib.matrixInit(sg.shape[0], sg.ctypes.data_as(C.POINTER(C.c_uint32)))

    def elem(v):
        buf = libc.calloc(1, 8)
        C.cast(buf, C.POINTER(C.c_uint64))[0] = int(v)
        return DcElement(sg.shape[0], C.cast(buf, C.POINTER(C.c_uint64)))
    for x, y, med, cost in g["lookups"][:40]:
        e1, e2, r = elem(x), elem(y), elem(0)
        c = lib.getCostAndMedian2D(C.byref(e1), C.byref(e2), C.byref(r), C.c_void_p(m))
        assert (int(r.element[0]), int(c)) == (med, cost)
        for e in (e1, e2, r):
            libc.free(C.cast(e.element, C.c_void_p))
    lib.matrixDestroy(C.c_void_p(m))


@pytest.mark.parametrize("dialect", [0, 1, 2])
@pytest.mark.parametrize("a,name", [(21, "pref-gap"), (21, "pref-sub"), (12, "random-sym"), (9, "random-asym"), (32, "pref-sub")])
def test_explicit_dialects_fuzz(oracle_mod, gpu_ctx, dialect, a, name):
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(4000 + 100 * dialect + a)
    sg = _sigma(name, a, rng)
    pairs = _pairs(rng, a, 120, 220)
    s = _rand_str(rng, 90, a)
    pairs += [(s, s.copy()), (s[:40], s[:40].copy())]
    # heavily ambiguous strings: many distinct elements -> the pair's table goes to global scratch
    pairs += [(_rand_str(rng, 150, a, amb=0.9), _rand_str(rng, 140, a, amb=0.9)) for _ in range(6)]
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    memo = gpu_ctx.memo(sg)
    r = metric_do_batch(gpu_ctx, memo, dialect, [p[0] for p in pairs], [p[1] for p in pairs])
    for k, (s1, s2) in enumerate(pairs):
        oc = oracle_mod.haskell_do(dialect, a, s1, s2, sigma=sg)
        gc = r.pair(k)
        assert gc[0] == oc[0], (k, len(s1), len(s2), gc[0], oc[0])
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), (k, len(s1), len(s2))
        assert int(r.final_offset[k]) == oc[5], (k, int(r.final_offset[k]), oc[5])
        assert int(r.cells[k]) == oc[4] or dialect != 0


def test_explicit_discrete_matrix_equals_metric_kernel(gpu_ctx):
    """The 0/1 matrix through the Sankoff table path must equal the closed-form discrete-metric kernel."""
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(12)
    a = 21
    sg = (1 - np.eye(a, dtype=np.uint32)).astype(np.uint32)
    pairs = _pairs(rng, a, 150, 250)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    memo = gpu_ctx.memo(sg)
    for dialect in (0, 1, 2):
        r1 = metric_do_batch(gpu_ctx, memo, dialect, [p[0] for p in pairs], [p[1] for p in pairs])
        r0 = metric_do_batch(gpu_ctx, a, dialect, [p[0] for p in pairs], [p[1] for p in pairs])
        assert np.array_equal(r0.cost, r1.cost)
        assert np.array_equal(r0.out_len, r1.out_len)
        assert np.array_equal(r0.final_offset, r1.final_offset)
        for k in range(len(pairs)):
            for x, y in zip(r0.pair(k)[1:], r1.pair(k)[1:]):
                assert np.array_equal(x, y), (dialect, k)


def test_multi_pass_schedule_equals_single_launch(gpu_ctx, monkeypatch):
    """PCGDO_EXPLICIT_MULTI_PASS=1 runs one Ukkonen offset per launch with device-side queues; same results."""
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(99)
    a = 21
    pairs = _pairs(rng, a, 200, 250)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    memo = gpu_ctx.memo(_sigma("pref-gap", a, rng))
    f, s = [p[0] for p in pairs], [p[1] for p in pairs]
    for dialect in (1, 2):
        for alpha in (a, memo):
            ref = metric_do_batch(gpu_ctx, alpha, dialect, f, s)
            monkeypatch.setenv("PCGDO_EXPLICIT_MULTI_PASS", "1")
            gpu_ctx.reload_env()
            r = metric_do_batch(gpu_ctx, alpha, dialect, f, s)
            monkeypatch.delenv("PCGDO_EXPLICIT_MULTI_PASS")
            gpu_ctx.reload_env()
            assert np.array_equal(ref.cost, r.cost) and np.array_equal(ref.final_offset, r.final_offset)
            assert np.array_equal(ref.out_len, r.out_len) and np.array_equal(ref.cells, r.cells)
            for k in range(len(pairs)):
                for x, y in zip(ref.pair(k)[1:], r.pair(k)[1:]):
                    assert np.array_equal(x, y), (dialect, k)


@pytest.mark.parametrize("a,kind", [(12, "discrete"), (21, "explicit"), (9, "l1"), (32, "explicit")])
def test_unboxed_full_matrix_dialect(oracle_mod, gpu_ctx, a, kind):
    """Dialect 3 = unboxedFullMatrixDO as coded (S2 cell over the full matrix, first column seeded from row 0,
    UnboxedFullMatrix.hs:86-90) against the oracle's restatement, all three overlap kinds.  PARITY UNPINNED upstream."""
    from pcg_b200.pairwise import metric_do_batch, S2_FULL
    rng = np.random.default_rng(600 + a)
    sg = _sigma("random-sym", a, rng) if kind == "explicit" else None
    pairs = _pairs(rng, a, 150, 90)
    # strings whose best alignment opens with deletions of the lesser string: the seeded first column decides
    for _ in range(20):
        s = _rand_str(rng, 40, a)
        pairs.append((np.concatenate([_rand_str(rng, int(rng.integers(1, 6)), a), s]), np.concatenate([s, _rand_str(rng, 8, a)])))
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=128, max_b=32, arena_words=0)
    memo = gpu_ctx.memo(sg, kind, a=a)
    r = metric_do_batch(gpu_ctx, memo, S2_FULL, [p[0] for p in pairs], [p[1] for p in pairs])
    differs = 0
    for k, (s1, s2) in enumerate(pairs):
        oc = oracle_mod.haskell_do(3, a, s1, s2, sigma=sg, metric_kind={"discrete": 0, "explicit": 1, "l1": 2}[kind])
        gc = r.pair(k)
        assert gc[0] == oc[0], (k, len(s1), len(s2), gc[0], oc[0])
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), (k, len(s1), len(s2))
        differs += oc[0] != oracle_mod.haskell_do(0, a, s1, s2, sigma=sg, metric_kind={"discrete": 0, "explicit": 1, "l1": 2}[kind])[0]
    assert int(r.cells[0]) == (len(pairs[0][0]) + 1) * (len(pairs[0][1]) + 1)


@pytest.mark.parametrize("a,kind,amb", [(21, "discrete", 0.0), (21, "pref-gap", 0.0), (21, "random-asym", 0.0), (12, "l1", 0.0),
                                        (4, "discrete", 0.5), (4, "random-sym", 0.5), (21, "pref-sub", 0.02)])
def test_batch_dictionary_equals_per_pair_tables(oracle_mod, gpu_ctx, monkeypatch, a, kind, amb):
    """When a whole batch draws its elements from at most 30 distinct ambiguity sets, one table over them — replicated
    per shared-memory bank, conflict-free — replaces the per-pair memo tables (DictDev).  Same costs, offsets, cells and
    alignments as the per-pair path (PCGDO_NO_DICT=1), every dialect; a = 4 with ambiguous elements brings gap-containing
    sets and medians that are exactly the gap (dialect 2 then keeps the per-pair tables), 2 % ambiguity over 21 symbols
    overflows the dictionary (more than 30 sets)."""
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(7 * a + len(kind))

    def rstr(n):
        x = (1 << rng.integers(0, a - 1, size=n)).astype(np.uint32)
        if amb:
            m = rng.random(n) < amb
            x = np.where(m, rng.integers(1, 1 << a, size=n), x).astype(np.uint32)
        return x
    pairs = []
    for _ in range(200):
        n1 = int(rng.integers(1, 200))
        s1 = rstr(n1)
        r = rng.random()
        if r < 0.4:
            s2 = s1.copy()
            mut = rng.random(n1) < 0.1
            s2 = np.where(mut, rstr(n1), s2).astype(np.uint32)[rng.random(n1) > 0.05]
            if len(s2) == 0:
                s2 = s1[:1].copy()
        else:
            s2 = rstr(int(rng.integers(1, 200)))
        pairs.append((s1, s2) if rng.random() < 0.5 else (s2, s1))
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=256, max_b=32, arena_words=0)
    if kind == "discrete":
        alpha = a
    elif kind == "l1":
        idx = np.arange(a)
        alpha = gpu_ctx.memo(np.abs(np.subtract.outer(idx, idx)).astype(np.uint32), structure="l1")
    else:
        alpha = gpu_ctx.memo(_sigma(kind, a, rng))
    f, s = [p[0] for p in pairs], [p[1] for p in pairs]
    for dialect in (0, 1, 2):
        r = metric_do_batch(gpu_ctx, alpha, dialect, f, s)
        monkeypatch.setenv("PCGDO_NO_DICT", "1")
        gpu_ctx.reload_env()
        ref = metric_do_batch(gpu_ctx, alpha, dialect, f, s)
        monkeypatch.delenv("PCGDO_NO_DICT")
        gpu_ctx.reload_env()
        assert np.array_equal(r.cost, ref.cost) and np.array_equal(r.final_offset, ref.final_offset), dialect
        assert np.array_equal(r.cells, ref.cells) and np.array_equal(r.out_len, ref.out_len), dialect
        for k in range(len(pairs)):
            for x, y in zip(r.pair(k)[1:], ref.pair(k)[1:]):
                assert np.array_equal(x, y), (dialect, k)
        if kind in ("discrete", "pref-gap") and a == 21:       # and against the restatement itself
            sig = None if kind == "discrete" else _sigma(kind, a, rng)
            for k in range(0, len(pairs), 9):
                oc = oracle_mod.haskell_do(dialect, a, pairs[k][0], pairs[k][1], sigma=sig)
                gc = r.pair(k)
                assert gc[0] == oc[0] and all(np.array_equal(x, y) for x, y in zip(gc[1:], oc[1:4])), (dialect, k)


@pytest.mark.parametrize("a", [21, 82])
def test_pipelined_host_path_for_large_alphabets(gpu_ctx, monkeypatch, a):
    """pcgdo_metric_batch_host / pcgdo_explicit_batch_host cut monotone batches into chunks (copy in / align / copy out on
    three streams): same results as the one-shot path, for one-word and multi-word elements."""
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(a)
    W = (a + 31) // 32

    def rstr(n):
        sym = rng.integers(0, a - 1, size=n)
        out = np.zeros((n, W), np.uint32)
        out[np.arange(n), sym // 32] = (1 << (sym % 32)).astype(np.uint32)
        return out if W > 1 else out.reshape(-1)
    pairs = [(rstr(int(rng.integers(1, 120))), rstr(int(rng.integers(1, 120)))) for _ in range(1500)]
    f, s = [p[0] for p in pairs], [p[1] for p in pairs]
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=128, max_b=32, arena_words=0)
    alpha = a if a <= 32 else gpu_ctx.memo(None, structure="discrete", a=a)
    monkeypatch.setenv("PCGDO_HOST_CHUNK_PAIRS", "1048576")
    gpu_ctx.reload_env()
    plain = metric_do_batch(gpu_ctx, alpha, 2, f, s)
    monkeypatch.setenv("PCGDO_HOST_CHUNK_PAIRS", "256")
    gpu_ctx.reload_env()
    piped = metric_do_batch(gpu_ctx, alpha, 2, f, s)
    monkeypatch.delenv("PCGDO_HOST_CHUNK_PAIRS")
    gpu_ctx.reload_env()
    for x, y in ((plain.cost, piped.cost), (plain.out_len, piped.out_len), (plain.cells, piped.cells),
                 (plain.final_offset, piped.final_offset)):
        assert np.array_equal(x, y)
    # the strings pair by pair: the one-shot path copies the whole output arrays back, unwritten padding behind the last
    # pair included (whatever the device buffer held), the pipelined one only what its chunks wrote
    for k in range(len(pairs)):
        assert all(np.array_equal(u, v) for u, v in zip(plain.pair(k)[1:], piped.pair(k)[1:])), k
