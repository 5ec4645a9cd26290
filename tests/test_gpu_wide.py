"""GPU parity: the additive (L1) closed form of the a > 8 path and alphabets of more than 32 symbols (multi-word
elements) through the per-pair table kernel, against the oracle's restatements (oracle/do_oracle.c, oracle/hs_wide.c) and —
end to end, post-order + re-rooting on the device forest — against eleven DAG costs the reference's own script tests
assert (test/TestSuite/ScriptTests.hs:205-266; tests/golden/custom_alphabet_scripts.json)."""
import numpy as np
import pytest

from tests.test_gpu_metric import _pairs

pytestmark = pytest.mark.gpu

_KIND = {"discrete": 0, "explicit": 1, "l1": 2}          # the oracle's metric_kind numbering


def _sigma(a, rng):
    m = rng.integers(1, 9, size=(a, a))
    m = np.minimum(m, m.T)
    np.fill_diagonal(m, 0)
    return m.astype(np.uint32)


@pytest.mark.parametrize("dialect", [0, 1, 2])
@pytest.mark.parametrize("a", [9, 21, 32])
def test_l1_dialects_fuzz(oracle_mod, gpu_ctx, dialect, a):
    """firstLinearNormPairwiseLogic (as coded upstream: the median of a disjoint pair drops the upper bound's symbol)."""
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(7000 + 10 * dialect + a)
    pairs = _pairs(rng, a, 140, 220)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    memo = gpu_ctx.memo(None, "l1", a=a)
    r = metric_do_batch(gpu_ctx, memo, dialect, [p[0] for p in pairs], [p[1] for p in pairs])
    for k, (s1, s2) in enumerate(pairs):
        oc = oracle_mod.haskell_do(dialect, a, s1, s2, metric_kind=2)
        gc = r.pair(k)
        assert gc[0] == oc[0], (k, len(s1), len(s2), gc[0], oc[0])
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), (k, len(s1), len(s2))
        assert int(r.final_offset[k]) == oc[5]


def test_structure_diagnosis(gpu_ctx):
    """PCGDO_METRIC_AUTO follows diagnoseTcm: |i-j| -> L1 (also for the 2 x 2 matrix that is 0/1 as well), 0/1 -> discrete."""
    from pcg_b200.backend import METRIC_DISCRETE, METRIC_EXPLICIT, METRIC_L1
    i = np.arange(12)
    assert gpu_ctx.memo(np.abs(np.subtract.outer(i, i)), "auto").structure == METRIC_L1
    assert gpu_ctx.memo(1 - np.eye(12, dtype=np.uint32), "auto").structure == METRIC_DISCRETE
    assert gpu_ctx.memo(1 - np.eye(2, dtype=np.uint32), "auto").structure == METRIC_L1
    assert gpu_ctx.memo(2 * (1 - np.eye(12, dtype=np.uint32)), "auto").structure == METRIC_EXPLICIT   # caller factors first


def _rand_wide(rng, n, a, amb=0.12, gapful=0.05):
    W = (a + 31) // 32
    out = np.zeros((n, W), np.uint32)
    for k in range(n):
        bits = {int(rng.integers(0, a - 1))}
        if rng.random() < amb:
            bits |= {int(b) for b in rng.integers(0, a - 1, size=int(rng.integers(1, 5)))}
        if rng.random() < gapful:
            bits.add(a - 1)
        for b in bits:
            out[k, b >> 5] |= np.uint32(1 << (b & 31))
    return out


def _wide_pairs(rng, a, count, max_len):
    out = []
    for _ in range(count):
        n1 = int(rng.integers(1, max_len))
        s1 = _rand_wide(rng, n1, a)
        if rng.random() < 0.6:                    # related strings: the Ukkonen band stays narrow
            s2 = s1.copy()
            mut = np.flatnonzero(rng.random(n1) < 0.08)
            if len(mut):
                s2[mut] = _rand_wide(rng, len(mut), a)
            s2 = s2[rng.random(n1) > 0.04]
            if len(s2) == 0:
                s2 = s1[:1].copy()
        else:
            s2 = _rand_wide(rng, max(1, n1 + int(rng.integers(-15, 15))), a)
        out.append((s1, s2) if rng.random() < 0.5 else (s2, s1))
    s = _rand_wide(rng, 40, a)
    out += [(s, s.copy()), (s[:7], s[:7].copy())]          # equal strings: measureCharacters falls through to EQ
    t = s.copy()
    t[20, -1] ^= np.uint32(1)                               # equal length, first difference in the most significant word
    if not t[20].any():                                     # (never an empty ambiguity set: not a valid element)
        t[20, -1] = np.uint32(2)
    out.append((s, t))
    return out


@pytest.mark.parametrize("dialect", [0, 1, 2])
@pytest.mark.parametrize("a,kind", [(33, "discrete"), (33, "explicit"), (82, "l1"), (82, "explicit"), (200, "discrete"),
                                    (257, "l1"), (140, "explicit")])
def test_wide_alphabet_dialects_fuzz(oracle_mod, gpu_ctx, dialect, a, kind):
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(9000 + 100 * dialect + a)
    sg = _sigma(a, rng) if kind == "explicit" else None
    pairs = _wide_pairs(rng, a, 60, 110)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=128, max_b=32, arena_words=0)
    memo = gpu_ctx.memo(sg, kind, a=a)
    assert memo.words == (a + 31) // 32
    r = metric_do_batch(gpu_ctx, memo, dialect, [p[0] for p in pairs], [p[1] for p in pairs])
    for k, (s1, s2) in enumerate(pairs):
        oc = oracle_mod.haskell_do_w(dialect, a, s1, s2, sigma=sg, metric_kind=_KIND[kind])
        gc = r.pair(k)
        assert gc[0] == oc[0], (k, len(s1), len(s2), gc[0], oc[0])
        for x, y in zip(gc[1:], oc[1:4]):
            assert np.array_equal(x, y), (k, len(s1), len(s2))
        assert int(r.final_offset[k]) == oc[5]


_CUSTOM = [("slashes", "L1-norm"), ("slashes", "discrete"), ("slashes", "2-1"), ("slashes", "hamming"),
           ("slashes", "levenshtein"), ("large-mix", "discrete"), ("huge-mix", "discrete"), ("huge-mix", "L1-norm"),
           ("huge-mix", "1-2"), ("huge-mix", "hamming"), ("huge-mix", "levenshtein"),
           # commented out upstream (their pre-order pass fails; the DO cost they record still holds):
           ("slashes", "1-2"), ("large-mix", "1-2"), ("large-mix", "2-1"), ("large-mix", "hamming"),
           ("large-mix", "levenshtein"), ("huge-mix", "2-1")]


@pytest.mark.parametrize("ds,name", _CUSTOM)
def test_wide_forest_reproduces_script_test_costs(gpu_ctx, ds, name):
    """The reference's script-test DAG costs on 82-, 256- and 433-symbol alphabets through the device: multi-word
    elements, metric chosen by the TCM diagnosis, S2 Ukkonen at every node, re-rooting."""
    import pcg_b200 as P
    from tests.helpers import custom_alphabet_golden
    a, children, leaves, cases = custom_alphabet_golden(ds)
    kind, sigma, weight, want = cases[name]
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=128, max_b=32, arena_words=0)
    memo = gpu_ctx.memo(sigma, "auto")
    assert memo.structure == {0: 1, 1: 0, 2: 2}[kind]            # oracle numbering -> PCGDO_METRIC_*
    f = P.Forest(gpu_ctx, children[None], len(leaves), [[s] for s in leaves], node_cap=128, metric=memo, dialect=2)
    f.enable_rerooting()
    assert weight * int(f.score_rerooted(None)[0]) == want
    f.close()
