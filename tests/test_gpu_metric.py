"""GPU parity: large-alphabet Haskell dialects (S1 full, S1 Ukkonen, S2 Ukkonen) over the discrete metric
against the oracle's restatement (oracle_haskell_do).  PARITY UNPINNED upstream for these paths."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_str(rng, n, a, amb=0.1, gapful=0.05):
    """Median strings over an a-symbol alphabet: single symbols, some ambiguity sets, a few containing the gap
    bit (but never the gap alone — the Haskell wrapper strips those before aligning)."""
    gap = 1 << (a - 1)
    x = (1 << rng.integers(0, a - 1, size=n)).astype(np.uint32)
    m = rng.random(n) < amb
    x = np.where(m, rng.integers(1, gap, size=n), x).astype(np.uint32)
    g = rng.random(n) < gapful
    return np.where(g, x | gap, x).astype(np.uint32)


def _pairs(rng, a, count, max_len):
    out = []
    for _ in range(count):
        n1 = int(rng.integers(1, max_len))
        r = rng.random()
        if r < 0.5:                               # related: Ukkonen band stays narrow
            s1 = _rand_str(rng, n1, a)
            s2 = s1.copy()
            mut = rng.random(n1) < 0.08
            s2 = np.where(mut, _rand_str(rng, n1, a), s2).astype(np.uint32)
            keep = rng.random(n1) > 0.04
            s2 = s2[keep]
            if len(s2) == 0:
                s2 = s1[:1].copy()
        elif r < 0.8:
            s1 = _rand_str(rng, n1, a)
            s2 = _rand_str(rng, max(1, n1 + int(rng.integers(-20, 20))), a)
        else:
            s1 = _rand_str(rng, n1, a)
            s2 = _rand_str(rng, int(rng.integers(1, max_len)), a)
        out.append((s1, s2) if rng.random() < 0.5 else (s2, s1))
    return out


@pytest.mark.parametrize("dialect", [0, 1, 2])
@pytest.mark.parametrize("a", [21, 5, 12])
def test_metric_dialects_fuzz(oracle_mod, gpu_ctx, dialect, a):
    import pcg_b200 as P
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(1000 + 10 * dialect + a)
    pairs = _pairs(rng, a, 160, 260)
    s = _rand_str(rng, 100, a)
    pairs += [(s, s.copy()), (s[:50], s[:50].copy())]
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    r = metric_do_batch(gpu_ctx, a, dialect, [p[0] for p in pairs], [p[1] for p in pairs])
    offs = set()
    for k, (s1, s2) in enumerate(pairs):
        oc = oracle_mod.haskell_do(dialect, a, s1, s2)
        gc = r.pair(k)
        assert gc[0] == oc[0], (k, len(s1), len(s2), gc[0], oc[0])
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), (k, len(s1), len(s2))
        assert int(r.final_offset[k]) == oc[5], (k, int(r.final_offset[k]), oc[5])
        offs.add(oc[5])
    if dialect:
        assert len(offs) >= 3      # full-matrix fall-backs and several band widths were exercised


def test_metric_matches_dense_table_path_on_dna(oracle_mod, gpu_ctx):
    """Cross-dialect sanity on a small alphabet: S1 full cost == C-linear cost whenever the C band is not binding
    (the reference's own cross-implementation predicate, app/2d-do-comparison.hs:119-128: same cost)."""
    import pcg_b200 as P
    from pcg_b200.pairwise import metric_do_batch
    from tests.helpers import tcm_matrix
    rng = np.random.default_rng(9)
    pairs = _pairs(rng, 5, 80, 120)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    r = metric_do_batch(gpu_ctx, 5, 0, [p[0] for p in pairs], [p[1] for p in pairs])
    tcm = gpu_ctx.tcm(tcm_matrix("discrete"))
    pp = P.PackedPairs.from_lists([p[0].astype(np.uint8) for p in pairs], [p[1].astype(np.uint8) for p in pairs])
    rl = P.align2d_batch(gpu_ctx, tcm, pp)
    assert np.array_equal(r.cost, rl.cost)


def test_register_kernel_equals_table_kernel(gpu_ctx, monkeypatch):
    """The discrete metric has two device implementations: the per-pair table kernel (default) and the older
    kernel that evaluates the overlap per cell in registers (PCGDO_METRIC_REGISTERS=1).  Same results required;
    PCGDO_NO_PACK16=1 additionally forces the table kernel's 32-bit fill."""
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(321)
    pairs = _pairs(rng, 21, 120, 240)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    f, s = [p[0] for p in pairs], [p[1] for p in pairs]
    for dialect in (0, 1, 2):
        ref = metric_do_batch(gpu_ctx, 21, dialect, f, s)
        for env in ("PCGDO_METRIC_REGISTERS", "PCGDO_NO_PACK16"):
            monkeypatch.setenv(env, "1")
            gpu_ctx.reload_env()
            r = metric_do_batch(gpu_ctx, 21, dialect, f, s)
            monkeypatch.delenv(env)
            gpu_ctx.reload_env()
            assert np.array_equal(ref.cost, r.cost) and np.array_equal(ref.final_offset, r.final_offset), (dialect, env)
            for k in range(len(pairs)):
                for x, y in zip(ref.pair(k)[1:], r.pair(k)[1:]):
                    assert np.array_equal(x, y), (dialect, env, k)
