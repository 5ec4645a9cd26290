"""GPU: candidate-tree turnover (pcgdo_forest_set_trees), the sweep's cell count, several GPUs in one process
(pcgdo_multi / pcgdo_mforest) and the short-string case of the compact host transfer."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.helpers import random_string, tcm_matrix

pytestmark = pytest.mark.gpu


def _forest_inputs(seed, n_trees, n_taxa, n_chars, lo=20, hi=90):
    from pcg_b200.postorder import random_binary_trees
    rng = np.random.default_rng(seed)
    children = random_binary_trees(n_trees, n_taxa, seed)
    leaves = [[random_string(rng, int(rng.integers(lo, hi)), amb=0.1) for _ in range(n_chars)] for _ in range(n_taxa)]
    return children, leaves


@pytest.mark.parametrize("reroot", [False, True])
def test_set_trees_equals_fresh_forest(gpu_ctx, reroot):
    """New topologies levelled on the device into the existing forest score exactly like a forest built from them."""
    import pcg_b200 as P
    n_trees, n_taxa, n_chars = 24, 40, 5
    ch_a, leaves = _forest_inputs(11, n_trees, n_taxa, n_chars)
    ch_b, _ = _forest_inputs(12, n_trees, n_taxa, n_chars)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=400, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(tcm_matrix("discrete"))
    score = (lambda f: f.score_rerooted(tcm)) if reroot else (lambda f: f.score(tcm))
    fresh = {}
    for name, ch in (("a", ch_a), ("b", ch_b)):
        f = P.Forest(gpu_ctx, ch, n_taxa, leaves, 400)
        if reroot:
            f.enable_rerooting()
        fresh[name] = score(f)
        f.close()
    f = P.Forest(gpu_ctx, ch_a, n_taxa, leaves, 400)
    if reroot:
        f.enable_rerooting()
    assert np.array_equal(score(f), fresh["a"])
    f.set_trees(ch_b)
    assert np.array_equal(score(f), fresh["b"])
    f.set_trees(ch_a)
    assert np.array_equal(score(f), fresh["a"])
    # a caterpillar (deepest possible tree: one level per internal node) next to balanced ones
    cat = ch_b.copy()
    cat[0, 0] = (0, 1)
    for x in range(1, n_taxa - 1):
        cat[0, x] = (x + 1, n_taxa + x - 1)
    f.set_trees(cat)
    got = score(f)
    g = P.Forest(gpu_ctx, cat, n_taxa, leaves, 400)
    if reroot:
        g.enable_rerooting()
    assert np.array_equal(got, score(g))
    assert f.levels == n_taxa - 1
    g.close()
    f.close()


def test_set_trees_rejects_what_is_not_a_tree(gpu_ctx):
    import pcg_b200 as P
    n_trees, n_taxa, n_chars = 4, 12, 2
    ch, leaves = _forest_inputs(3, n_trees, n_taxa, n_chars)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=200, max_b=32, arena_words=0)
    f = P.Forest(gpu_ctx, ch, n_taxa, leaves, 200)
    for mutate in ("shared_child", "bad_id", "cycle"):
        bad = ch.copy()
        if mutate == "shared_child":
            bad[2, 5, 0] = bad[2, 4, 0]
        elif mutate == "bad_id":
            bad[1, 3, 1] = 2 * n_taxa + 5
        else:
            bad[3, 0] = (0, n_taxa + n_taxa - 2)      # the first internal node takes the root as a child
        with pytest.raises(P.BackendError):
            f.set_trees(bad)
    f.set_trees(ch)                                   # still usable afterwards
    tcm = gpu_ctx.tcm(tcm_matrix("discrete"))
    g = P.Forest(gpu_ctx, ch, n_taxa, leaves, 200)
    assert np.array_equal(f.score(tcm), g.score(tcm))
    f.close(); g.close()


def test_forest_cells_equals_reference_band_cells(oracle_mod, gpu_ctx):
    import pcg_b200 as P
    n_trees, n_taxa, n_chars = 6, 30, 3
    ch, leaves = _forest_inputs(21, n_trees, n_taxa, n_chars, 30, 200)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=600, max_b=32, arena_words=0)
    sig = tcm_matrix("l1")
    tcm = gpu_ctx.tcm(sig)
    f = P.Forest(gpu_ctx, ch, n_taxa, leaves, 600)
    f.score(tcm)
    flat = np.concatenate([s for taxon in leaves for s in taxon])
    ln = np.array([len(s) for taxon in leaves for s in taxon], np.uint32)
    off = np.concatenate([[0], np.cumsum(ln.astype(np.uint64))[:-1]]).astype(np.uint64)
    _, cost, cells = oracle_mod.cpu_bench_postorder(sig, ch, n_taxa, n_chars, flat, off, ln, 0, n_trees * n_chars, 4, "port")
    assert f.cells() == int(cells.sum())
    f.close()


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    return [[0]] + ([[0, 1]] if n >= 2 else []) + ([list(range(n))] if n > 2 else [])


@pytest.mark.parametrize("n_chars", [7, 1])
def test_multi_forest_in_one_process(gpu_ctx, n_chars):
    """pcgdo_mforest over 1, 2 and all visible devices, every reduction mode, post-order and re-rooted, tree turnover:
    the per-tree costs equal the single-device forest's.  n_chars = 1 exercises the split by trees."""
    import pcg_b200 as P
    from pcg_b200.backend import REDUCE_HOST, REDUCE_PEER, REDUCE_NCCL
    n_trees, n_taxa = 16, 24
    ch, leaves = _forest_inputs(31, n_trees, n_taxa, n_chars)
    ch2, _ = _forest_inputs(32, n_trees, n_taxa, n_chars)
    sig = tcm_matrix("discrete")
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=400, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    want = {}
    for name, c in (("a", ch), ("b", ch2)):
        f = P.Forest(gpu_ctx, c, n_taxa, leaves, 400)
        want[name] = f.score(tcm)
        f.enable_rerooting()
        want[name + "r"] = f.score_rerooted(tcm)
        f.close()
    for devs in _device_lists():
        m = P.Multi(devs)
        m.configure(8, 16, 400, 32)
        mt = m.tcm(sig)
        mf = m.forest(ch, n_taxa, leaves, 400)
        modes = [REDUCE_HOST] + ([REDUCE_PEER, REDUCE_NCCL] if len(devs) > 1 else [])
        for mode in modes:
            try:
                got, wall, sweep, red = m.score(mt, mf, reduce_mode=mode)
            except P.BackendError as ex:
                if mode == REDUCE_HOST:
                    raise
                pytest.skip(f"reduce mode {mode} unavailable here: {ex}")
            assert np.array_equal(got, want["a"]), (devs, mode)
            assert wall > 0 and sweep > 0
        m.set_trees(mf, ch2)
        assert np.array_equal(m.score(mt, mf)[0], want["b"]), devs
        m.enable_rerooting(mf)
        assert np.array_equal(m.score(mt, mf, rerooted=True)[0], want["br"]), devs
        blocks = [m.block(mf, i) for i in range(len(devs))]
        assert blocks[0][0] == 0 and blocks[-1][1] == (n_trees if blocks[0][2] else n_chars)
        m.destroy_forest(mf)
        m.close()


def test_multi_pair_batch_in_one_process(oracle_mod, gpu_ctx):
    import pcg_b200 as P
    from pcg_b200.backend import Batch
    from tests.helpers import random_pair
    rng = np.random.default_rng(8)
    pairs = [random_pair(rng, 150) for _ in range(3001)]
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    sig = tcm_matrix("l1")
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=200, max_b=32, arena_words=0)
    one = P.align2d_batch(gpu_ctx, gpu_ctx.tcm(sig), pp)
    for devs in _device_lists():
        m = P.Multi(devs)
        m.configure(8, 32, 200, 32)
        mt = m.tcm(sig)
        cost = np.zeros(pp.n, np.int32); out_len = np.zeros(pp.n, np.uint32)
        g, s1, s2 = (np.zeros(pp.out_bytes, np.uint8) for _ in range(3))
        p = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
        b = Batch(pp.n, p(pp.elems), pp.elems.nbytes, p(pp.off1), p(pp.off2), p(pp.len1), p(pp.len2), p(pp.out_off),
                  pp.out_bytes, p(g), p(s1), p(s2), p(out_len), p(cost), None, None)
        m.align2d_batch_host(mt, b)
        assert np.array_equal(cost, one.cost) and np.array_equal(out_len, one.out_len)
        for k in range(pp.n):
            o, n = int(pp.out_off[k]), int(out_len[k])
            assert np.array_equal(g[o:o + n], one.gapped[o:o + n]) and np.array_equal(s1[o:o + n], one.slot1[o:o + n]) \
                and np.array_equal(s2[o:o + n], one.slot2[o:o + n]), k
        m.close()


@pytest.mark.parametrize("chunk", [256, 1024])
def test_compact_transfer_of_very_short_strings(oracle_mod, gpu_ctx, monkeypatch, chunk):
    """Pairs of 1 .. 8 elements: their packed 16-byte units would not fit the chunk's reserved region (round4(len1 +
    len2) per pair), so such chunks pack in 4-byte units; nothing may spill into the next chunk or past out_bytes."""
    import pcg_b200 as P
    monkeypatch.setenv("PCGDO_HOST_CHUNK_PAIRS", str(chunk))
    gpu_ctx.reload_env()
    sig = tcm_matrix("discrete")
    o = oracle_mod.Oracle(sig)
    rng = np.random.default_rng(4)
    pairs = [(random_string(rng, int(rng.integers(1, 9))), random_string(rng, int(rng.integers(1, 9)))) for _ in range(6000)]
    pairs[4000:4400] = [(random_string(rng, int(rng.integers(40, 90))), random_string(rng, int(rng.integers(40, 90)))) for _ in range(400)]
    pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=128, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    plain = P.align2d_batch(gpu_ctx, tcm, pp)
    guard = 4096
    n = pp.n
    cost = np.zeros(n, np.int32); out_len = np.zeros(n, np.uint32); actual = np.zeros(n, np.uint64)
    bufs = [np.full(pp.out_bytes + guard, 0xAB, np.uint8) for _ in range(3)]
    from pcg_b200.backend import Batch
    p = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    b = Batch(n, p(pp.elems), pp.elems.nbytes, p(pp.off1), p(pp.off2), p(pp.len1), p(pp.len2), p(pp.out_off), pp.out_bytes,
              p(bufs[0]), p(bufs[1]), p(bufs[2]), p(out_len), p(cost), None, p(actual))
    gpu_ctx.check(gpu_ctx.lib.pcgdo_align2d_batch_host(gpu_ctx.h, tcm.h, C.byref(b)), "pcgdo_align2d_batch_host")
    for buf in bufs:
        assert (buf[pp.out_bytes:] == 0xAB).all()                # nothing written past the caller's arrays
    assert np.array_equal(cost, plain.cost) and np.array_equal(out_len, plain.out_len)
    order = np.argsort(actual)
    ends = actual[order] + out_len[order]
    assert (ends[:-1] <= actual[order][1:]).all()                # ranges are disjoint
    for k in range(n):
        oa, o0, ln = int(actual[k]), int(pp.out_off[k]), int(out_len[k])
        assert np.array_equal(bufs[0][oa:oa + ln], plain.gapped[o0:o0 + ln]), k
        assert np.array_equal(bufs[1][oa:oa + ln], plain.slot1[o0:o0 + ln]) and np.array_equal(bufs[2][oa:oa + ln], plain.slot2[o0:o0 + ln]), k
    for k in (0, 1, 255, 256, 4100, 5999):
        oc = o.align2d(*pairs[k])
        assert oc[0] == cost[k] and np.array_equal(bufs[0][int(actual[k]):int(actual[k]) + int(out_len[k])], oc[1])
