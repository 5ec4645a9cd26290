"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, bit-exact."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests.helpers import bench_strings_golden, crc_u8, random_pair, random_string, tcm_matrix

pytestmark = pytest.mark.gpu


def _check_batch(o, ctx, tcm, pairs, want_cells=True):
    import pcg_b200 as P
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    r = P.align2d_batch(ctx, tcm, pp, want_cells=want_cells)
    for k, (s1, s2) in enumerate(pairs):
        oc = o.align2d(s1, s2)
        gc = r.pair(pp, k)
        assert gc[0] == oc[0], (k, len(s1), len(s2), gc[0], oc[0])
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), (k, len(s1), len(s2))
        if want_cells:
            assert int(r.cells[k]) == oc[4], (k, len(s1), len(s2))
    return r


@pytest.mark.parametrize("name", ["discrete", "l1", "1:2", "2:1", "weird"])
def test_fuzz_against_oracle(oracle_mod, gpu_ctx, name):
    sig = tcm_matrix(name)
    o = oracle_mod.Oracle(sig)
    gpu_ctx.configure(max_len=640, max_b=32)
    tcm = gpu_ctx.tcm(sig)
    rng = np.random.default_rng(101)
    pairs = [random_pair(rng, 600) for _ in range(300)]
    _check_batch(o, gpu_ctx, tcm, pairs)


def test_edge_shapes(oracle_mod, gpu_ctx):
    sig = tcm_matrix("discrete")
    o = oracle_mod.Oracle(sig)
    gpu_ctx.configure(max_len=640, max_b=32)
    tcm = gpu_ctx.tcm(sig)
    rng = np.random.default_rng(7)
    pairs = []
    for n1, n2 in [(1, 1), (1, 2), (2, 1), (1, 300), (300, 1), (5, 5), (16, 17), (31, 33), (64, 64), (65, 63),
                   (99, 100), (128, 127), (200, 133), (200, 134), (150, 100), (151, 100), (500, 499),
                   (512, 512), (600, 401), (600, 399), (57, 58), (59, 52), (61, 61)]:
        pairs.append((random_string(rng, n1), random_string(rng, n2)))
    s = random_string(rng, 257)
    pairs.append((s, s.copy()))                       # identical inputs
    t = s.copy(); t[200] ^= 3
    pairs.append((s, t)); pairs.append((t, s))         # equal length, lexicographic tie-break
    g = np.full(40, 16, np.uint8)
    pairs.append((g, random_string(rng, 44)))          # all-gap elements as input symbols
    pairs.append((np.full(90, 31, np.uint8), np.full(70, 31, np.uint8)))   # fully ambiguous
    _check_batch(o, gpu_ctx, tcm, pairs)


@pytest.mark.parametrize("a,name", [(3, "discrete"), (4, "l1"), (6, "discrete"), (7, "l1"), (8, "discrete")])
def test_other_alphabets(oracle_mod, gpu_ctx, a, name):
    sig = tcm_matrix(name, a)
    o = oracle_mod.Oracle(sig)
    gpu_ctx.configure(max_len=320, max_b=32)
    tcm = gpu_ctx.tcm(sig)
    rng = np.random.default_rng(a)
    pairs = []
    for _ in range(60):
        s1, s2 = random_pair(rng, 300, a=a)
        pairs.append((s1, s2))
    _check_batch(o, gpu_ctx, tcm, pairs)


def test_arena_recycling_and_small_batches(oracle_mod, gpu_ctx):
    """Tiny arena: warps must trace/emit and restart several times within one hand-out of pairs."""
    sig = tcm_matrix("l1")
    o = oracle_mod.Oracle(sig)
    gpu_ctx.configure(pairs_per_warp=32, warps_per_cta=4, max_len=320, max_b=32, arena_words=40000)
    tcm = gpu_ctx.tcm(sig)
    rng = np.random.default_rng(9)
    pairs = [random_pair(rng, 300) for _ in range(200)]
    _check_batch(o, gpu_ctx, tcm, pairs)
    gpu_ctx.configure(pairs_per_warp=3, warps_per_cta=16, max_len=320, max_b=32, arena_words=0)
    _check_batch(o, gpu_ctx, tcm, pairs[:50])
    gpu_ctx.configure(pairs_per_warp=32)


def test_golden_nodes_through_gpu(oracle_mod, gpu_ctx):
    """The reference's chel golden test, post-order nodes, through foreignPairwiseDO on the GPU."""
    import pcg_b200 as P
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "chel_golden.json")))
    gpu_ctx.configure(max_len=640, max_b=32)
    tcm = gpu_ctx.tcm(tcm_matrix("discrete"))
    nodes = g["nodes"]
    ids = [n["id"] for n in nodes if len(n["children"]) == 2 and n["id"] not in g["rerooted_nodes"]]
    c1 = [np.array(nodes[sorted(nodes[i]["children"])[0]]["context"], np.uint8) for i in ids]
    c2 = [np.array(nodes[sorted(nodes[i]["children"])[1]]["context"], np.uint8) for i in ids]
    got = P.foreign_pairwise_do_batch(gpu_ctx, tcm, c1, c2)
    assert len(ids) == 14
    for i, (cost, ctx) in zip(ids, got):
        assert cost == nodes[i]["local_cost"], i
        assert ctx.tolist() == nodes[i]["context"], i


def test_reference_compatible_align2d_symbol(oracle_mod, gpu_ctx):
    """Call `align2d` exactly as DO/Pairwise/FFI.hsc does: malloc'ed alignIO_t, payload right-aligned."""
    import pcg_b200.backend as be
    lib = be.load()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    sig = tcm_matrix("discrete")
    o = oracle_mod.Oracle(sig)
    cm = be.CostMatrices2D()
    lib.setUp2dCostMtx(C.byref(cm), sig.ctypes.data_as(C.POINTER(C.c_uint32)), 5, 0)

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

    rng = np.random.default_rng(23)
    for _ in range(6):
        s1, s2 = random_pair(rng, 200)
        cap = len(s1) + len(s2) + 2
        a1, a2, g, u = aio(s1, cap), aio(s2, cap), aio([], cap), aio([], cap)
        cost = lib.align2d(C.byref(a1), C.byref(a2), C.byref(g), C.byref(u), C.byref(cm), 0, 1, 0)
        oc = o.align2d(s1, s2)
        assert cost == oc[0]
        assert take(g) == oc[1].tolist() and take(a1) == oc[2].tolist() and take(a2) == oc[3].tolist()
        libc.free(C.cast(u.character, C.c_void_p))


def test_config2_shape_properties_and_sample(oracle_mod, gpu_ctx):
    """BASELINE config 2 shape (1 kb uniform DNA pairs, L1 TCM, C band) at a size the oracle cannot
    sweep: size-independent properties on every pair + the oracle on a sample."""
    import pcg_b200 as P
    sig = tcm_matrix("l1")
    o = oracle_mod.Oracle(sig)
    gpu_ctx.configure(pairs_per_warp=32, warps_per_cta=16, max_len=1024, max_b=8, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    n, L = 8192, 1000
    rng = np.random.Generator(np.random.PCG64(20260101))
    strings = (1 << rng.integers(0, 4, size=(2 * n, L))).astype(np.uint8)
    pp = P.PackedPairs.uniform(strings)
    r = P.align2d_batch(gpu_ctx, tcm, pp, want_cells=True)
    cost_tab = o.cost
    gap = o.gap
    assert (r.cells == r.cells[0]).all() and int(r.cells[0]) == o.band_cells(L + 1, L + 1)[0]
    for k in range(0, n, 37):
        c, g, s1, s2 = r.pair(pp, k)
        a, b = strings[2 * k], strings[2 * k + 1]
        # each slot holds one input with gaps (0) inserted ...
        first_long = a.tolist() >= b.tolist()         # equal lengths: lexicographic (c_alignment_interface.c:59-70)
        x, y = (s1, s2) if first_long else (s2, s1)          # x = aligned lesser, y = aligned longer
        lesser, longer = (b, a) if first_long else (a, b)
        assert np.array_equal(x[x != 0], lesser) and np.array_equal(y[y != 0], longer)
        assert not ((x == 0) & (y == 0)).any()
        # ... the reported cost is the sum of the column costs, the median the table's
        xe = np.where(x == 0, gap, x).astype(np.int64)
        ye = np.where(y == 0, gap, y).astype(np.int64)
        assert int(cost_tab[ye, xe].sum()) == c
        assert np.array_equal(o.median[xe, ye].astype(np.uint8), g)
    for k in range(0, n, 257):
        oc = o.align2d(strings[2 * k], strings[2 * k + 1])
        gc = r.pair(pp, k)
        assert gc[0] == oc[0]
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y)
    # commutativity: swapping the inputs swaps the slots, nothing else
    sw = strings.reshape(n, 2, L)[:, ::-1, :].reshape(2 * n, L)
    pp2 = P.PackedPairs.uniform(np.ascontiguousarray(sw[:512]))
    r2 = P.align2d_batch(gpu_ctx, tcm, pp2)
    for k in range(256):
        c, g, s1, s2 = r.pair(pp, k)
        c2, g2, t1, t2 = r2.pair(pp2, k)
        assert c == c2 and np.array_equal(g, g2) and np.array_equal(s1, t2) and np.array_equal(s2, t1)


def test_general_geometry_kernel(oracle_mod, gpu_ctx):
    """Pairs the warp kernels cannot hold (forced by a tiny max_len / max_b, and genuinely wide ones)
    take the one-CTA-per-pair kernel and must give the same bits."""
    sig = tcm_matrix("l1")
    o = oracle_mod.Oracle(sig)
    tcm = gpu_ctx.tcm(sig)
    rng = np.random.default_rng(77)
    pairs = [random_pair(rng, 400) for _ in range(40)]
    pairs.append((random_string(rng, 2300), random_string(rng, 900)))     # full matrix, 3200 diagonals
    pairs.append((random_string(rng, 1500), random_string(rng, 1400)))
    pairs.append((random_string(rng, 700), random_string(rng, 1)))
    gpu_ctx.configure(max_len=64, max_b=8)          # almost everything overflows the staging window
    _check_batch(o, gpu_ctx, tcm, pairs)
    gpu_ctx.configure(max_len=2400, max_b=8)        # wide kernel disabled: B > 8 pairs go to the general kernel
    _check_batch(o, gpu_ctx, tcm, pairs)
    gpu_ctx.configure(max_len=2400, max_b=32)       # everything that fits a warp stays on the warp kernels
    _check_batch(o, gpu_ctx, tcm, pairs)


def test_bench_strings_dna_all_pairs(oracle_mod, gpu_ctx):
    """BASELINE config 1 shape: every unordered pair of 20 strings with lengths 8 ... 4608 (half of them
    ambiguous), discrete TCM — synthetic stand-ins with the reference file's lengths."""
    sig = tcm_matrix("discrete")
    o = oracle_mod.Oracle(sig)
    tcm = gpu_ctx.tcm(sig)
    rng = np.random.default_rng(20)
    lens = [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 9, 18, 36, 72, 144, 288, 576, 1152, 2304, 4608]
    strs = [random_string(rng, n, amb=0.0 if k < 10 else 0.5) for k, n in enumerate(lens)]
    pairs = [(strs[i], strs[j]) for i in range(20) for j in range(i + 1, 20)]
    gpu_ctx.configure(max_len=4700, max_b=32)
    _check_batch(o, gpu_ctx, tcm, pairs, want_cells=True)


def test_pipelined_host_entry_point(oracle_mod, gpu_ctx):
    """More pairs than one host chunk: copies of chunk c+1 / c-1 overlap the alignment of chunk c; results
    must equal the single-shot path, including when a chunk contains pairs for the wide / general kernels."""
    import pcg_b200 as P
    sig = tcm_matrix("l1")
    o = oracle_mod.Oracle(sig)
    tcm = gpu_ctx.tcm(sig)
    rng = np.random.default_rng(5)
    pairs = [random_pair(rng, 120) for _ in range(70000)]
    pairs[100] = (random_string(rng, 700), random_string(rng, 400))       # wide kernel (B = 16/32)
    pairs[69000] = (random_string(rng, 500), random_string(rng, 30))
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=720, max_b=32, arena_words=0)
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    r = P.align2d_batch(gpu_ctx, tcm, pp)
    for k in list(range(0, 70000, 997)) + [100, 69000, 65535, 65536, 69999]:
        oc = o.align2d(*pairs[k])
        gc = r.pair(pp, k)
        assert gc[0] == oc[0], k
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), k
    # a pair only the general kernel can hold forces the plain path for the whole batch
    pairs[5] = (random_string(rng, 2000), random_string(rng, 600))
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    r = P.align2d_batch(gpu_ctx, tcm, pp)
    for k in [5, 100, 6, 69999]:
        oc = o.align2d(*pairs[k])
        gc = r.pair(pp, k)
        assert gc[0] == oc[0], k
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), k


@pytest.mark.parametrize("chunk,pinned", [(0, False), (8192, False), (0, True), (8192, True)])
def test_compact_transfer_equals_plain(oracle_mod, gpu_ctx, monkeypatch, chunk, pinned):
    """pcgdo_batch.out_off_actual: outputs packed per pipelined chunk (warps claim ranges from a per-chunk cursor once
    a pair's alignment length is known) must be, pair by pair, the bytes the plain transfer returns — with the default
    chunk (2 chunks) and with 9 chunks cycling through the 3 staging slots; with pinned output arrays the kernel writes
    them directly over PCIe (zero-copy, growing chunks); a batch that takes the plain path (one chunk, or a pair only
    the general kernel can hold) reports out_off unchanged."""
    import pcg_b200 as P
    if chunk:
        monkeypatch.setenv("PCGDO_HOST_CHUNK_PAIRS", str(chunk))
    if pinned:
        monkeypatch.setenv("PCGDO_ZERO_COPY", "1")       # the direct-store experiment (opt-in: measured slower than staging)
    gpu_ctx.reload_env()
    sig = tcm_matrix("l1")
    o = oracle_mod.Oracle(sig)
    tcm = gpu_ctx.tcm(sig)
    rng = np.random.default_rng(15)
    pairs = [random_pair(rng, 120) for _ in range(70000)]
    pairs[100] = (random_string(rng, 600), random_string(rng, 420))       # band of 284 diagonals: wide kernel (B = 16)
    pairs[300] = (random_string(rng, 1), random_string(rng, 1))
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=720, max_b=32, arena_words=0)
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    plain = P.align2d_batch(gpu_ctx, tcm, pp)
    packed = P.align2d_batch(gpu_ctx, tcm, pp, compact=True, pinned=pinned)
    assert np.array_equal(plain.cost, packed.cost) and np.array_equal(plain.out_len, packed.out_len)
    # it did pack: the ranges are disjoint and a chunk's last byte lies well before its reserved capacity ends
    order = np.argsort(packed.out_off)
    ends = packed.out_off[order] + ((packed.out_len[order].astype(np.uint64) + 3) & ~np.uint64(3))
    assert (ends[:-1] <= packed.out_off[order][1:]).all()
    assert int(ends[-1]) < int(pp.out_off[-1])
    for k in range(70000):
        a, b = plain.pair(pp, k), packed.pair(pp, k)
        assert all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:])), k
    for k in (0, 100, 300, 8191, 8192, 65535, 65536, 69999):
        oc = o.align2d(*pairs[k])
        gc = packed.pair(pp, k)
        assert gc[0] == oc[0] and all(np.array_equal(x, y) for x, y in zip(gc[1:], oc[1:4])), k
    small = P.PackedPairs.from_lists([p[0] for p in pairs[:500]], [p[1] for p in pairs[:500]])
    r = P.align2d_batch(gpu_ctx, tcm, small, compact=True, pinned=pinned)   # one chunk: plain path, offsets as reserved
    assert np.array_equal(r.out_off, small.out_off)
    assert all(np.array_equal(x, y) for x, y in zip(r.pair(small, 499)[1:], plain.pair(pp, 499)[1:]))
    # pairs only the general-geometry kernel can hold: aligned on their own after the pipeline, their strings placed
    # behind the packed bytes of their chunk (the zero-copy experiment redoes the batch on the plain path instead)
    pairs[5] = (random_string(rng, 2000), random_string(rng, 600))
    pairs[66000] = (random_string(rng, 1800), random_string(rng, 700))
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    r = P.align2d_batch(gpu_ctx, tcm, pp, compact=True, pinned=pinned)
    if pinned:
        assert np.array_equal(r.out_off, pp.out_off)
    else:
        assert (r.out_off != pp.out_off).any() and int(r.out_off.max()) < int(pp.out_off[-1])      # still packed
        order = np.argsort(r.out_off)
        ends = r.out_off[order] + ((r.out_len[order].astype(np.uint64) + 3) & ~np.uint64(3))
        assert (ends[:-1] <= r.out_off[order][1:]).all()          # disjoint ranges, stragglers included
    for k in (5, 6, 4, 8191, 8192, 65999, 66000, 66001, 69999):
        oc = o.align2d(*pairs[k])
        gc = r.pair(pp, k)
        assert gc[0] == oc[0] and all(np.array_equal(x, y) for x, y in zip(gc[1:], oc[1:4])), k


@pytest.mark.parametrize("tcm_file", ["dna-discrete.tcm", "dna-L1-norm.tcm", "dna-pref-gap.tcm", "dna-pref-sub.tcm"])
def test_bench_strings_golden_through_gpu(gpu_ctx, tcm_file):
    """BASELINE config 1 on the device against the committed outputs of the compiled reference (no oracle
    in the loop): 190 pairs, lengths 8 ... 4608, IUPAC ambiguity, all four DNA TCM files."""
    import pcg_b200 as P
    enc, per_tcm = bench_strings_golden()
    sigma, recs = per_tcm[tcm_file]
    tcm = gpu_ctx.tcm(sigma)
    gpu_ctx.configure(max_len=4700, max_b=32)
    firsts, seconds = [], []
    for i, j, *_ in recs:
        a, b = (enc[i], enc[j]) if len(enc[i]) >= len(enc[j]) else (enc[j], enc[i])
        firsts.append(a); seconds.append(b)
    pp = P.PackedPairs.from_lists(firsts, seconds)
    r = P.align2d_batch(gpu_ctx, tcm, pp)
    for k, (i, j, cost, n, cg, c1, c2) in enumerate(recs):
        gc = r.pair(pp, k)
        assert (gc[0], len(gc[1]), crc_u8(gc[1]), crc_u8(gc[2]), crc_u8(gc[3])) == (cost, n, cg, c1, c2), (i, j)


def test_packed_fill_equals_32bit_fill(oracle_mod, gpu_ctx, monkeypatch):
    """The packed 16-bit fill (default when the TCM leaves room in its window) against the 32-bit fill
    (PCGDO_NO_PACK16=1 at TCM creation), and a TCM whose costs force the window to overflow -> fallback."""
    import pcg_b200 as P
    rng = np.random.default_rng(77)
    pairs = [random_pair(rng, 700) for _ in range(300)]
    pairs += [(random_string(rng, 1500), random_string(rng, 1450)), (random_string(rng, 3000, amb=0.0), random_string(rng, 2900, amb=0.0))]
    gpu_ctx.configure(max_len=3100, max_b=32)
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    for name in ("l1", "weird"):
        sig = tcm_matrix(name)
        r16 = P.align2d_batch(gpu_ctx, gpu_ctx.tcm(sig), pp)
        monkeypatch.setenv("PCGDO_NO_PACK16", "1")
        gpu_ctx.reload_env()
        r32 = P.align2d_batch(gpu_ctx, gpu_ctx.tcm(sig), pp)
        monkeypatch.delenv("PCGDO_NO_PACK16")
        gpu_ctx.reload_env()
        assert np.array_equal(r16.cost, r32.cost), name
        assert np.array_equal(r16.out_len, r32.out_len), name
        assert np.array_equal(r16.gapped, r32.gapped) and np.array_equal(r16.slot1, r32.slot1), name
    # large indel costs: 4 * 60 per step, the finite spread of long unrelated strings leaves the 15-bit window
    # indel 60: statically outside the packed window; indel 20: inside it statically, but the sweep of long
    # unrelated strings overflows it dynamically -> redone by the 32-bit fill
    for indel in (60, 20):
        sig = tcm_matrix("discrete") * 3
        sig[-1, :-1] = indel
        sig[:-1, -1] = indel
        o = oracle_mod.Oracle(sig)
        _check_batch(o, gpu_ctx, gpu_ctx.tcm(sig), pairs[:60] + pairs[-2:])


def _grouped_pairs(rng, n_groups, lo=330, hi=1040, per_group=4):
    """Groups of four consecutive pairs of about the same size (what the four-pairs-per-warp kernel sweeps at once),
    with ragged members: unequal operands, related strings, ambiguity, the odd group that does not qualify."""
    from tests.helpers import random_string
    pairs = []
    for gi in range(n_groups):
        base = int(rng.integers(lo, hi))
        for _ in range(per_group):
            n1 = max(1, base + int(rng.integers(-25, 25)))
            n2 = max(1, n1 - int(rng.integers(0, 40))) if rng.random() < 0.7 else n1
            if gi % 11 == 10 and rng.random() < 0.3:
                n2 = int(rng.integers(1, max(2, lo // 2)))     # a member that breaks up its group
            s1 = random_string(rng, n1, 5, 0.15)
            if rng.random() < 0.5:
                s2 = s1[:n2].copy()
                mut = rng.random(n2) < 0.08
                s2 = np.where(mut, random_string(rng, n2, 5, 0.15), s2).astype(np.uint8)
            else:
                s2 = random_string(rng, n2, 5, 0.15)
            pairs.append((s1, s2) if rng.random() < 0.5 else (s2, s1))
    return pairs


@pytest.mark.parametrize("shape", ["4x8x28", "8x4x32", "6x5x28", "5x6x28", "mixed"])
@pytest.mark.parametrize("name", ["discrete", "l1", "1:2", "2:1", "weird"])
def test_several_pairs_per_warp_kernels(oracle_mod, gpu_ctx, monkeypatch, name, shape):
    """do_linear_multi_kernel — four pairs per warp for bands of 169..223 diagonals, five for 141..167, six for 129..139, eight for 65..127 — against the CPU
    side (the compiled reference when present) and against the one-pair-per-warp kernels on the same batch."""
    import zlib
    import pcg_b200 as P
    sig = tcm_matrix(name)
    rng = np.random.default_rng(4242)
    if shape == "4x8x28":
        pairs = _grouped_pairs(rng, 150)
    elif shape == "8x4x32":
        pairs = _grouped_pairs(rng, 120, lo=30, hi=290, per_group=8)
    elif shape == "6x5x28":
        pairs = _grouped_pairs(rng, 150, lo=270, hi=360, per_group=6)
    elif shape == "5x6x28":
        pairs = _grouped_pairs(rng, 150, lo=400, hi=640, per_group=5)
    else:
        pairs = _grouped_pairs(rng, 60) + _grouped_pairs(rng, 60, lo=30, hi=290, per_group=8) + \
            _grouped_pairs(rng, 60, lo=270, hi=360, per_group=6) + _grouped_pairs(rng, 60, lo=400, hi=640, per_group=5) + \
            _grouped_pairs(rng, 20, lo=900, hi=1400)
        rng.shuffle(pairs)
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    gpu_ctx.configure(pairs_per_warp=32, warps_per_cta=32, max_len=1500 if shape == "mixed" else 1100, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    all_shapes = shape in ("8x4x32", "mixed")                  # the 8-pairs shape is off by default (measured: no gain)
    if all_shapes:
        monkeypatch.setenv("PCGDO_QUAD_FLAGS", "32")
        gpu_ctx.reload_env()
    try:
        r = P.align2d_batch(gpu_ctx, tcm, pp, want_cells=True)
        n_launch_quad = gpu_ctx.last_kernel_ms()[1]
    finally:
        if all_shapes:
            monkeypatch.delenv("PCGDO_QUAD_FLAGS")
        gpu_ctx.reload_env()
    monkeypatch.setenv("PCGDO_NO_QUAD", "1")
    gpu_ctx.reload_env()
    try:
        r1 = P.align2d_batch(gpu_ctx, tcm, pp, want_cells=True)
        n_launch_single = gpu_ctx.last_kernel_ms()[1]
    finally:
        monkeypatch.delenv("PCGDO_NO_QUAD")
        gpu_ctx.reload_env()
    assert n_launch_quad == n_launch_single + (4 if all_shapes else 3)        # the extra kernels did run
    assert np.array_equal(r.cost, r1.cost) and np.array_equal(r.out_len, r1.out_len)
    assert np.array_equal(r.cells, r1.cells)
    kind = "reference" if oracle_mod.Reference.available() else "port"
    cost, out_len, crc = oracle_mod.cpu_align2d_full(sig, pp, max(1, min(32, len(os.sched_getaffinity(0)))), kind)
    assert np.array_equal(r.cost, cost) and np.array_equal(r.out_len, out_len)
    for k in range(pp.n):
        o, ln = int(pp.out_off[k]), int(r.out_len[k])
        got = (zlib.crc32(r.gapped[o:o + ln].tobytes()), zlib.crc32(r.slot1[o:o + ln].tobytes()),
               zlib.crc32(r.slot2[o:o + ln].tobytes()))
        assert got == tuple(int(x) for x in crc[k]), (k, len(pairs[k][0]), len(pairs[k][1]))
