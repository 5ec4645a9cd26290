"""GPU parity at BASELINE.json's sizes, DIRECTLY against the compiled, unmodified reference (oracle/_ref) where it
exists: the reference runs every pair on host threads (oracle/cpu_bench.c) and hands back, per pair, the cost, the
alignment length and the CRC-32 of each of its three output strings; the device's strings must hash to the same
values.  The workloads are bench.py's generators (same seeds as the bench lines), cut to what the CPU finishes in
seconds.  Where upstream is Haskell (config 4) the restatement stands in, with every overlap answered by the compiled
reference tcm-memo."""
import os
import types
import zlib

import numpy as np
import pytest

import bench as B
from tests.helpers import tcm_matrix

pytestmark = pytest.mark.gpu

THREADS = max(1, min(32, len(os.sched_getaffinity(0))))


def _crc_rows(r, pp, n):
    """CRC-32 of the three output strings of pairs 0..n-1 of a BatchResult."""
    out = np.zeros((n, 3), np.uint32)
    off = pp.out_off if r.out_off is None else r.out_off
    for k in range(n):
        o, ln = int(off[k]), int(r.out_len[k])
        out[k] = (zlib.crc32(r.gapped[o:o + ln].tobytes()), zlib.crc32(r.slot1[o:o + ln].tobytes()),
                  zlib.crc32(r.slot2[o:o + ln].tobytes()))
    return out


def _sub(pp, n):
    return types.SimpleNamespace(n=n, elems=pp.elems, off1=pp.off1[:n].copy(), off2=pp.off2[:n].copy(),
                                 len1=pp.len1[:n].copy(), len2=pp.len2[:n].copy(), out_off=pp.out_off[:n].copy(),
                                 out_bytes=int(pp.out_off[n - 1]) + int((int(pp.len1[n - 1]) + int(pp.len2[n - 1]) + 3) & ~3) + 16)


@pytest.mark.parametrize("name,variant", [("discrete", 1), ("l1", 1), ("discrete", 0), ("l1", 0)])
def test_config3_affine_5kb_against_compiled_reference(oracle_mod, gpu_ctx, name, variant):
    """>= 256 pairs of BASELINE config 3 (lengths U[4750, 5250], 10 % IUPAC sets, gap-open 3) through
    pcgdo_align2d_affine_batch_host: cost, length and all three strings equal to `align2dAffine` of the compiled
    reference — the patched build for variant 1, the unmodified one for variant 0 (EXT/c_alignment_interface.c:180-348).
    The 12/16-diagonal classes and the > 512-diagonal row-wise fill carry most of these pairs."""
    import pcg_b200 as P
    so = oracle_mod.REF_PATCHED_SO if variant else oracle_mod.REF_SO
    kind = "reference" if os.path.exists(so) else "port"
    n = 288
    pp = _sub(B.config3_pairs(n), n)
    sig = tcm_matrix(name)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=5250, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig, gap_open=3, affine_variant=variant)
    r = P.align2d_batch(gpu_ctx, tcm, pp)
    cost, out_len, crc = oracle_mod.cpu_align2d_full(sig, pp, THREADS, kind, gap_open=3, affine_variant=variant)
    assert np.array_equal(r.cost, cost), np.flatnonzero(r.cost != cost)[:8]
    assert np.array_equal(r.out_len, out_len)
    assert np.array_equal(_crc_rows(r, pp, n), crc)
    wq = np.abs(pp.len1.astype(np.int64) - pp.len2.astype(np.int64)) + 8 + 41
    assert (wq > 512).any() and (wq <= 128).any() and ((wq > 256) & (wq <= 384)).any()      # every band class present


def test_config3_affine_pipelined_host_path(oracle_mod, gpu_ctx, monkeypatch):
    """The chunked copy / align / copy pipeline of the affine host entry point returns what the one-shot path does."""
    import pcg_b200 as P
    n = 640
    pp = _sub(B.config3_pairs(n), n)
    sig = tcm_matrix("discrete")
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=5250, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig, gap_open=3, affine_variant=1)
    monkeypatch.setenv("PCGDO_HOST_CHUNK_PAIRS", "1048576")
    gpu_ctx.reload_env()
    plain = P.align2d_batch(gpu_ctx, tcm, pp)
    monkeypatch.setenv("PCGDO_HOST_CHUNK_PAIRS", "256")
    gpu_ctx.reload_env()
    piped = P.align2d_batch(gpu_ctx, tcm, pp)
    assert np.array_equal(plain.cost, piped.cost) and np.array_equal(plain.out_len, piped.out_len)
    assert np.array_equal(_crc_rows(plain, pp, n), _crc_rows(piped, pp, n))


def test_config2_1kb_against_compiled_reference(oracle_mod, gpu_ctx):
    """20 000 pairs of the config-2 workload (bench seed) against `align2d` of the compiled reference: cost, length,
    three strings — through the pipelined host entry point with compact transfer, as bench.py's e2e leg runs it."""
    import pcg_b200 as P
    kind = "reference" if oracle_mod.Reference.available() else "port"
    n = 20000
    pp = P.PackedPairs.uniform(B.make_strings(n))
    sig = B.l1_tcm()
    gpu_ctx.configure(pairs_per_warp=32, warps_per_cta=32, max_len=1024, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    os.environ["PCGDO_HOST_CHUNK_PAIRS"] = "4096"
    try:
        gpu_ctx.reload_env()
        r = P.align2d_batch(gpu_ctx, tcm, pp, compact=True)
    finally:
        del os.environ["PCGDO_HOST_CHUNK_PAIRS"]
        gpu_ctx.reload_env()
    cost, out_len, crc = oracle_mod.cpu_align2d_full(sig, pp, THREADS, kind)
    assert np.array_equal(r.cost, cost) and np.array_equal(r.out_len, out_len)
    assert np.array_equal(_crc_rows(r, pp, n), crc)


def _fuzz_batch(rng, n, max_len):
    from tests.helpers import random_pair
    pairs = [random_pair(rng, max_len) for _ in range(n)]
    return pairs


def test_linear_fuzz_against_compiled_reference(oracle_mod, gpu_ctx):
    """The fuzz set of test_gpu_linear (ragged lengths, ambiguity, all five TCMs) against the COMPILED REFERENCE
    itself, closing the chain GPU = restatement = reference on one and the same inputs."""
    import pcg_b200 as P
    if not oracle_mod.Reference.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(2024)
    pairs = _fuzz_batch(rng, 1500, 700)
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=720, max_b=32, arena_words=0)
    for name in ("discrete", "l1", "1:2", "2:1", "weird"):
        sig = tcm_matrix(name)
        r = P.align2d_batch(gpu_ctx, gpu_ctx.tcm(sig), pp)
        cost, out_len, crc = oracle_mod.cpu_align2d_full(sig, pp, THREADS, "reference")
        assert np.array_equal(r.cost, cost), name
        assert np.array_equal(r.out_len, out_len), name
        assert np.array_equal(_crc_rows(r, pp, pp.n), crc), name


@pytest.mark.parametrize("kind", ["discrete", "pref-gap"])
def test_config4_amino_acids_400(oracle_mod, gpu_ctx, kind):
    """>= 2000 pairs of BASELINE config 4 (400 residues, a = 21, bench seed), dialect S2 (unboxedUkkonenFullSpaceDO):
    cost, final Ukkonen offset, band cells, length and the three strings against the restated dialect — for the
    explicit matrix with EVERY overlap answered by the compiled reference tcm-memo (getCostAndMedian2D)."""
    from pcg_b200.pairwise import metric_do_batch
    A, L4, n = 21, 400, 2048
    strings = B.config4_strings(n)
    elems = np.ascontiguousarray(strings.reshape(-1))
    k = np.arange(n, dtype=np.uint64)
    off1 = k * np.uint64(2 * L4)
    off2 = off1 + np.uint64(L4)
    ln = np.full(n, L4, np.uint32)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=L4, max_b=32, arena_words=0)
    first, second = [strings[2 * i] for i in range(n)], [strings[2 * i + 1] for i in range(n)]
    if kind == "discrete":
        r = metric_do_batch(gpu_ctx, A, 2, first, second)
        cost, cells, out_len, crc, fo = oracle_mod.cpu_haskell_full(2, A, elems, off1, ln, off2, ln, THREADS)
    else:
        sig = B.pref_gap_tcm(A)
        r = metric_do_batch(gpu_ctx, gpu_ctx.memo(sig), 2, first, second)
        cost, cells, out_len, crc, fo = oracle_mod.cpu_haskell_full(
            2, A, elems, off1, ln, off2, ln, THREADS, sigma=sig, through_reference_memo=oracle_mod.ReferenceMemo.available())
    assert np.array_equal(r.cost, cost)
    assert np.array_equal(r.final_offset, fo)
    # the device counts the cells of the FINAL band (the restatement's counter also includes the refills of the
    # doubling schedule): rows 0..m, columns j with j - i in [-o, (n - m) + o] clipped to 0..n; o = -1: full matrix
    def final_band(o, m=L4, nn=L4):
        if o < 0:
            return (m + 1) * (nn + 1)
        i = np.arange(m + 1)
        return int((np.minimum(nn, i + (nn - m) + o) - np.maximum(0, i - o) + 1).clip(min=0).sum())
    assert np.array_equal(r.cells, np.array([final_band(int(o)) for o in fo], np.uint64))
    assert (cells >= r.cells).all()
    assert np.array_equal(r.out_len, out_len)
    for i in range(n):
        _, m, a1, a2 = r.pair(i)
        assert (zlib.crc32(m.tobytes()), zlib.crc32(a1.tobytes()), zlib.crc32(a2.tobytes())) == tuple(int(x) for x in crc[i]), i


def test_config5_forest_512_taxa_against_compiled_reference(oracle_mod, gpu_ctx):
    """BASELINE config 5's shape — 512 taxa, 256-element leaves, the bench's trees and leaves — on a column set of
    4 trees x 4 characters: EVERY internal node's local cost, total cost and ungapped median (8176 nodes) against a
    post-order loop over `align2d` of the compiled reference."""
    import pcg_b200 as P
    kind = "reference" if oracle_mod.Reference.available() else "port"
    n_trees, n_taxa, n_chars, leaf_len = 4, 512, 4, 256
    children, leaves = B.po_inputs(n_trees, n_taxa, n_chars, leaf_len)
    gpu_ctx.configure(pairs_per_warp=32, warps_per_cta=32, max_len=1024, max_b=32, arena_words=0)
    sig = B.discrete_tcm()
    tcm = gpu_ctx.tcm(sig)
    forest = P.Forest(gpu_ctx, children, n_taxa, [[leaves[v, c] for c in range(n_chars)] for v in range(n_taxa)], 1024)
    tree_cost = forest.score(tcm)
    flat = np.ascontiguousarray(leaves.reshape(-1))
    off = np.arange(n_taxa * n_chars, dtype=np.uint64) * np.uint64(leaf_len)
    ln = np.full(n_taxa * n_chars, leaf_len, np.uint32)
    col_cost, tot, loc, crc = oracle_mod.cpu_postorder_full(sig, children, n_taxa, n_chars, flat, off, ln, 0,
                                                           n_trees * n_chars, THREADS, kind)
    assert np.array_equal(tree_cost, col_cost.reshape(n_trees, n_chars).sum(axis=1))
    for t in range(n_trees):
        for c in range(n_chars):
            col = t * n_chars + c
            for x in range(n_taxa - 1):
                med, lc, tc = forest.node(t, c, n_taxa + x)
                assert (lc, tc, zlib.crc32(med.tobytes())) == (int(loc[col, x]), int(tot[col, x]), int(crc[col, x])), (t, c, x)
    forest.close()


@pytest.mark.parametrize("dialect", [0, 1, 3])
def test_s1_dialects_agree_with_compiled_align2d_where_its_band_is_not_binding(oracle_mod, gpu_ctx, dialect):
    """The reference's own cross-implementation predicate (app/2d-do-comparison.hs:119-128: the same cost from every
    pairwise DO implementation) as the pin upstream offers for naiveDO / ukkonenDO / unboxedFullMatrixDO: on DNA
    inputs whose lengths make `align2d` fill the FULL matrix (longer >= 1.5 x shorter: its band is not binding), the
    device's S1 / full-matrix dialects must cost exactly what the compiled C aligner does."""
    from pcg_b200.pairwise import metric_do_batch
    import pcg_b200 as P
    if not oracle_mod.Reference.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(90 + dialect)
    pairs = []
    for _ in range(300):
        n2 = int(rng.integers(1, 120))
        n1 = int(rng.integers((3 * n2 + 1) // 2 + 1, 2 * n2 + 40))
        pairs.append(((1 << rng.integers(0, 4, size=n1)).astype(np.uint8), (1 << rng.integers(0, 4, size=n2)).astype(np.uint8)))
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    for name in ("discrete", "1:2", "l1", "2:1", "weird"):
        sig = tcm_matrix(name)
        cost, _, _ = oracle_mod.cpu_align2d_full(sig, pp, THREADS, "reference")
        # the Haskell dialects take (lesser, longer) median strings; element = ambiguity set over 5 symbols
        alpha = 5 if name == "discrete" else gpu_ctx.memo(sig, structure="explicit")
        r = metric_do_batch(gpu_ctx, alpha, dialect, [p[1].astype(np.uint32) for p in pairs],
                            [p[0].astype(np.uint32) for p in pairs])
        assert np.array_equal(r.cost, cost), (name, np.flatnonzero(r.cost != cost)[:8])
