"""GPU parity: the post-order level scheduler (pcgdo_forest_*) against the oracle's post-order and the
reference's golden tree."""
import json
import os

import numpy as np
import pytest

from tests.helpers import random_string, tcm_matrix

pytestmark = pytest.mark.gpu


def _oracle_postorder(o, children, n_taxa, leaves):
    """leaves: per taxon element string. Returns {node: (context, local, total)} via the oracle's
    restatement of foreignPairwiseDO (DO/Internal.hs:127-140 for the totals)."""
    res = {v: (np.stack([leaves[v]] * 3, 1), 0, 0) for v in range(n_taxa)}
    pending = {n_taxa + x: tuple(int(c) for c in children[x]) for x in range(n_taxa - 1)}
    while pending:
        for n in [n for n, (l, r) in pending.items() if l in res and r in res]:
            l, r = pending.pop(n)
            cost, ctx, _ = o.pairwise_do(res[l][0], res[r][0])
            res[n] = (ctx, cost, cost + res[l][2] + res[r][2])
    return res


def test_forest_matches_oracle_postorder(oracle_mod, gpu_ctx):
    import pcg_b200 as P
    from pcg_b200.postorder import random_binary_trees
    sig = tcm_matrix("discrete")
    o = oracle_mod.Oracle(sig)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=400, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    n_trees, n_taxa, n_chars = 5, 14, 3
    rng = np.random.default_rng(31)
    children = random_binary_trees(n_trees, n_taxa, seed=4)
    leaves = [[random_string(rng, int(rng.integers(40, 90)), amb=0.05) for _ in range(n_chars)] for _ in range(n_taxa)]
    for v in range(n_taxa):          # leaves are ungapped strings: no pure-gap elements
        for c in range(n_chars):
            leaves[v][c][leaves[v][c] == 16] = 1
    f = P.Forest(gpu_ctx, children, n_taxa, leaves, node_cap=512)
    costs = f.score(tcm)
    assert f.alignments == n_trees * n_chars * (n_taxa - 1) and f.levels >= 4
    for t in range(n_trees):
        tree_cost = 0
        for c in range(n_chars):
            ref = _oracle_postorder(o, children[t], n_taxa, [leaves[v][c] for v in range(n_taxa)])
            is_child = set(int(x) for x in children[t].reshape(-1))
            for x in range(n_taxa - 1):
                node = n_taxa + x
                med, local, total = f.node(t, c, node)
                ctx, rl, rt = ref[node]
                assert local == rl and total == rt, (t, c, node)
                assert np.array_equal(med, ctx[ctx[:, 0] != o.gap][:, 0]), (t, c, node)
                if node not in is_child:
                    tree_cost += rt
        assert int(costs[t]) == tree_cost, t
    f.close()


def test_forest_on_reference_golden_tree(oracle_mod, gpu_ctx):
    """chel golden test: the 14 post-order nodes the reference pins — local costs and (ungapped) medians
    from the device scheduler, full contexts from the level-batched wrapper."""
    import pcg_b200 as P
    from pcg_b200.postorder import postorder_contexts
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "chel_golden.json")))
    nodes = g["nodes"]
    leaf_ids = [n["id"] for n in nodes if not n["children"]]
    int_ids = [n["id"] for n in nodes if n["children"]]
    n_taxa = len(leaf_ids)
    assert n_taxa == 17 and len(int_ids) == 16
    to_mine = {gid: k for k, gid in enumerate(leaf_ids)}
    to_mine.update({gid: n_taxa + k for k, gid in enumerate(int_ids)})
    children = np.zeros((1, n_taxa - 1, 2), np.int32)
    for k, gid in enumerate(int_ids):
        l, r = sorted(nodes[gid]["children"])            # lower node index = left (Postorder.hs:98-123)
        children[0, k] = (to_mine[l], to_mine[r])
    gap = 16
    leaves = []
    for gid in leaf_ids:
        c = np.array(nodes[gid]["context"], np.uint8)
        assert (c[:, 0] == c[:, 1]).all() and (c[:, 0] == c[:, 2]).all()
        leaves.append([c[c[:, 0] != gap][:, 0].copy()])
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=400, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(tcm_matrix("discrete"))
    f = P.Forest(gpu_ctx, children, n_taxa, leaves, node_cap=512)
    f.score(tcm)
    checked = 0
    for gid in int_ids:
        if gid in g["rerooted_nodes"]:
            continue
        med, local, total = f.node(0, 0, to_mine[gid])
        want = np.array(nodes[gid]["context"], np.uint8)
        assert local == nodes[gid]["local_cost"], gid
        assert total == nodes[gid]["total_cost"], gid
        assert np.array_equal(med, want[want[:, 0] != gap][:, 0]), gid
        checked += 1
    assert checked == 14
    f.close()
    # full contexts (with re-inserted gaps and left/right fields), level by level through foreignPairwiseDO;
    # the golden leaves are fed as they are stored (gapped contexts)
    leaf_ctx = [np.array(nodes[gid]["context"], np.uint8)[:, 0] for gid in leaf_ids]
    ctxs, local, total = postorder_contexts(gpu_ctx, tcm, children[0], n_taxa, leaf_ctx)
    for gid in int_ids:
        if gid in g["rerooted_nodes"]:
            continue
        assert local[to_mine[gid]] == nodes[gid]["local_cost"]
        assert ctxs[to_mine[gid]].tolist() == nodes[gid]["context"], gid


def test_forest_reports_node_overflow(gpu_ctx):
    import pcg_b200 as P
    from pcg_b200.postorder import random_binary_trees
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=400, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(tcm_matrix("discrete"))
    rng = np.random.default_rng(2)
    n_taxa = 8
    leaves = [[random_string(rng, 60, amb=0.0)] for _ in range(n_taxa)]
    f = P.Forest(gpu_ctx, random_binary_trees(2, n_taxa, 1), n_taxa, leaves, node_cap=60)
    with pytest.raises(P.BackendError):
        f.score(tcm)      # medians of unrelated strings grow beyond 60 elements: must fail loudly
    f.close()


def test_forest_rerooting_matches_oracle(oracle_mod, gpu_ctx):
    """Directed subtree data + edge costs on the device against the restatement of the re-rooting pass."""
    import pcg_b200 as P
    from pcg_b200.postorder import random_binary_trees
    sig = tcm_matrix("discrete")
    o = oracle_mod.Oracle(sig)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=600, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    n_trees, n_taxa, n_chars = 6, 13, 2
    rng = np.random.default_rng(77)
    children = random_binary_trees(n_trees, n_taxa, seed=11)
    base = [random_string(rng, 70, amb=0.0) for _ in range(n_chars)]
    leaves = []
    for v in range(n_taxa):          # related strings: mutated copies of a common ancestor, ragged lengths
        row = []
        for c in range(n_chars):
            s = base[c].copy()
            mut = rng.random(len(s)) < 0.25
            s = np.where(mut, random_string(rng, len(s), amb=0.1), s).astype(np.uint8)
            s = s[rng.random(len(s)) > 0.1]
            s[s == 16] = 2
            row.append(s)
        leaves.append(row)
    f = P.Forest(gpu_ctx, children, n_taxa, leaves, node_cap=768)
    plain = f.score(tcm)
    f.enable_rerooting()
    again = f.score(tcm)             # the post-order alone still works on the grown storage
    assert np.array_equal(plain, again)
    costs = f.score_rerooted(tcm)
    for t in range(n_trees):
        want = 0
        for c in range(n_chars):
            best, edges, ec, el = oracle_mod.reroot(o, children[t], [leaves[v][c] for v in range(n_taxa)])
            uv, tot, loc = f.edges(t, c)
            got = {(int(a), int(b)): (int(x), int(y)) for (a, b), x, y in zip(uv, tot, loc)}
            ref = {(int(a), int(b)): (int(x), int(y)) for (a, b), x, y in zip(edges, ec, el)}
            assert {tuple(sorted(k)): v for k, v in got.items()} == {tuple(sorted(k)): v for k, v in ref.items()}, (t, c)
            want += best
        assert int(costs[t]) == want, t
        assert int(costs[t]) <= int(plain[t])
    f.close()


def test_forest_rerooting_on_reference_golden_tree(gpu_ctx):
    """chel golden test through the device: re-rooted total 336, local cost 49 on the single minimal edge."""
    import pcg_b200 as P
    from tests.helpers import chel_golden_tree
    g, children, leaves, _ = chel_golden_tree()
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=400, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(tcm_matrix("discrete"))
    f = P.Forest(gpu_ctx, children, 17, [[s] for s in leaves], node_cap=512)
    f.enable_rerooting()
    costs = f.score_rerooted(tcm)
    assert int(costs[0]) == 336 == g["nodes"][0]["total_cost"]
    uv, tot, loc = f.edges(0, 0)
    k = int(np.argmin(tot))
    assert (tot == 336).sum() == 1 and int(loc[k]) == 49 == g["nodes"][0]["local_cost"]
    f.close()


def test_forest_wagner_edge_trials_match_oracle(oracle_mod, gpu_ctx):
    """iterativeBuild.tryEdge on the device: cost increase of attaching a new taxon to every edge of every tree."""
    import pcg_b200 as P
    from pcg_b200.postorder import random_binary_trees
    sig = tcm_matrix("1:2")
    o = oracle_mod.Oracle(sig)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=600, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    n_trees, n_taxa, n_chars = 4, 10, 3
    rng = np.random.default_rng(123)
    children = random_binary_trees(n_trees, n_taxa, seed=5)
    base = [random_string(rng, 60, amb=0.0) for _ in range(n_chars)]

    def variant(c):
        s = base[c].copy()
        mut = rng.random(len(s)) < 0.3
        s = np.where(mut, random_string(rng, len(s), amb=0.1), s).astype(np.uint8)
        s = s[rng.random(len(s)) > 0.1]
        s[s == 16] = 4
        return s
    leaves = [[variant(c) for c in range(n_chars)] for _ in range(n_taxa)]
    new_leaf = [variant(c) for c in range(n_chars)]
    f = P.Forest(gpu_ctx, children, n_taxa, leaves, node_cap=768)
    f.enable_rerooting(keep_edge_data=True)
    costs = f.score_rerooted(tcm)
    delta = f.try_leaf(tcm, new_leaf)
    assert delta.shape == (n_trees, 2 * n_taxa - 3)
    for t in range(n_trees):
        want_cost = 0
        want_delta = np.zeros(2 * n_taxa - 3, np.int64)
        order = None
        for c in range(n_chars):
            best, edges, ec, el, et = oracle_mod.reroot(o, children[t], [leaves[v][c] for v in range(n_taxa)], new_leaf[c])
            uv, tot, loc = f.edges(t, c)
            dev_pos = {tuple(sorted((int(a), int(b)))): k for k, (a, b) in enumerate(uv)}
            for (a, b), x, tr in zip(edges, ec, et):
                k = dev_pos[tuple(sorted((a, b)))]
                assert int(tot[k]) == int(x)
                want_delta[k] += int(tr)
            want_cost += best
        assert int(costs[t]) == want_cost
        assert np.array_equal(delta[t], want_delta), t
    f.close()


@pytest.mark.parametrize("kind", ["discrete", "explicit"])
def test_forest_large_alphabet_postorder(oracle_mod, gpu_ctx, kind):
    """Post-order over 21-symbol characters (32-bit elements, unboxedUkkonenFullSpaceDO at every node) against a
    host-side post-order that calls the restated dialect pair by pair."""
    import pcg_b200 as P
    from pcg_b200.postorder import random_binary_trees
    from tests.test_gpu_metric import _rand_str
    a, gap = 21, 1 << 20
    rng = np.random.default_rng(2024)
    sigma = None
    if kind == "explicit":
        sigma = 2 * (1 - np.eye(a, dtype=np.uint32))
        sigma[-1, :-1] = 1
        sigma[:-1, -1] = 1
        sigma = sigma.astype(np.uint32)
    n_trees, n_taxa, n_chars = 3, 9, 2
    children = random_binary_trees(n_trees, n_taxa, seed=8)
    base = [_rand_str(rng, 70, a, amb=0.0, gapful=0.0) for _ in range(n_chars)]

    def variant(c):
        s = base[c].copy()
        mut = rng.random(len(s)) < 0.2
        s = np.where(mut, _rand_str(rng, len(s), a, amb=0.1, gapful=0.0), s).astype(np.uint32)
        return s[rng.random(len(s)) > 0.08]
    leaves = [[variant(c) for c in range(n_chars)] for _ in range(n_taxa)]
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
    metric = a if sigma is None else gpu_ctx.memo(sigma)
    f = P.Forest(gpu_ctx, children, n_taxa, leaves, node_cap=256, metric=metric, dialect=2)
    costs = f.score(None)
    for t in range(n_trees):
        tree_cost = 0
        for c in range(n_chars):
            res = {v: (leaves[v][c], 0) for v in range(n_taxa)}
            pending = {n_taxa + x: tuple(int(z) for z in children[t][x]) for x in range(n_taxa - 1)}
            while pending:
                for n in [n for n, (l, r) in pending.items() if l in res and r in res]:
                    l, r = pending.pop(n)
                    cost, med, _, _, _, _ = oracle_mod.haskell_do(2, a, res[l][0], res[r][0], sigma=sigma)
                    res[n] = (med[med != gap], cost + res[l][1] + res[r][1])
                    got_med, got_local, got_total = f.node(t, c, n)
                    assert got_local == cost and got_total == res[n][1], (t, c, n)
                    assert np.array_equal(got_med, res[n][0]), (t, c, n)
            is_child = set(int(z) for z in children[t].reshape(-1))
            root = [n_taxa + x for x in range(n_taxa - 1) if n_taxa + x not in is_child][0]
            tree_cost += res[root][1]
        assert int(costs[t]) == tree_cost
    # re-rooting and edge trials run on the same scheduler: structural checks (the pass itself is pinned on DNA)
    f.enable_rerooting(keep_edge_data=True)
    rer = f.score_rerooted(None)
    assert (rer <= costs).all()
    for t in range(n_trees):
        for c in range(n_chars):
            uv, tot, loc = f.edges(t, c)
            is_child = set(int(z) for z in children[t].reshape(-1))
            root = [n_taxa + x for x in range(n_taxa - 1) if n_taxa + x not in is_child][0]
            assert int(tot[0]) == f.node(t, c, root)[2]
    delta = f.try_leaf(None, [variant(c) for c in range(n_chars)])
    assert delta.shape == (n_trees, 2 * n_taxa - 3) and (delta >= 0).all()
    f.close()


@pytest.mark.parametrize("name", ["1-2", "2-1"])
def test_large_alphabet_forest_reproduces_script_test_costs(gpu_ctx, name):
    """The reference's script-test KATs (ScriptTests.hs:199-204: DAG cost 1948 / 1242) through the device: 22 protein
    taxa, a = 21, explicit matrix through the per-pair memo table, S2 Ukkonen at every node, re-rooting."""
    import pcg_b200 as P
    from tests.helpers import invertebrates_golden
    children, leaves, cases = invertebrates_golden()
    sigma, want = cases[name]
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=400, max_b=32, arena_words=0)
    memo = gpu_ctx.memo(sigma)
    f = P.Forest(gpu_ctx, children[None], len(leaves), [[s] for s in leaves], node_cap=512, metric=memo, dialect=2)
    plain = f.score(None)
    f.enable_rerooting()
    assert int(f.score_rerooted(None)[0]) == want < int(plain[0])
    f.close()


def test_forest_reproduces_missing_values_script_cost(gpu_ctx):
    """scriptCheckCost 28 (ScriptTests.hs) through the device: dense g4t4 TCM, '?' elements, re-rooting."""
    import pcg_b200 as P
    from tests.helpers import missing_values_golden
    children, leaves, sigma, want = missing_values_golden()
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=64, max_b=32, arena_words=0)
    f = P.Forest(gpu_ctx, children[None], len(leaves), [[s] for s in leaves], node_cap=64)
    f.enable_rerooting()
    assert int(f.score_rerooted(gpu_ctx.tcm(sigma))[0]) == want
    f.close()


def test_forest_reproduces_multiblock_missing_script_cost(gpu_ctx):
    """scriptCheckCost 45 (ScriptTests.hs:184-186) through the device: two blocks with their own dense TCMs, each with
    one taxon Missing (PCGDO_MISSING leaves), re-rooting; 16 + 29."""
    import pcg_b200 as P
    from tests.helpers import multiblock_missing_golden
    children, blocks, want = multiblock_missing_golden()
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=64, max_b=32, arena_words=0)
    costs = []
    for leaves, sigma in blocks:
        f = P.Forest(gpu_ctx, children[None], len(leaves), [[s] for s in leaves], node_cap=64)
        f.enable_rerooting()
        costs.append(int(f.score_rerooted(gpu_ctx.tcm(sigma))[0]))
        f.close()
    assert costs == [16, 29] and sum(costs) == want


@pytest.mark.parametrize("wide", [False, True])
def test_forest_missing_and_empty_leaves_match_oracle(oracle_mod, gpu_ctx, wide):
    """Random trees whose leaves are, here and there, Missing or empty (all-gap taxa ungap to nothing): post-order and
    re-rooted costs, every edge's cost and the surviving node strings against the oracle's restatement of
    handleMissingCharacter / the empty branches of directOptimization — dense 8-bit forest and 32-bit (discrete, a = 12)."""
    import pcg_b200 as P
    from tests.helpers import tcm_matrix, random_string
    from tests.test_gpu_metric import _rand_str
    rng = np.random.default_rng(2024 + int(wide))
    n_taxa, n_trees, n_chars = 9, 6, 5
    from pcg_b200.postorder import random_binary_trees
    children = random_binary_trees(n_trees, n_taxa, 77)
    a = 12
    leaves = []
    for v in range(n_taxa):
        row = []
        for c in range(n_chars):
            r = rng.random()
            if r < 0.25:
                row.append(None)
            elif r < 0.4:
                row.append(np.zeros(0, np.uint32 if wide else np.uint8))
            else:
                n = int(rng.integers(5, 50))
                row.append(_rand_str(rng, n, a) if wide else random_string(rng, n))
        leaves.append(row)
    for c in range(n_chars):                 # one character with a single surviving taxon, one with none
        if c == 3:
            for v in range(1, n_taxa):
                leaves[v][c] = None
        if c == 4:
            for v in range(n_taxa):
                leaves[v][c] = None
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=128, max_b=32, arena_words=0)
    if wide:
        f = P.Forest(gpu_ctx, children, n_taxa, leaves, node_cap=128, metric=a, dialect=2)
        tcm = None
    else:
        sig = tcm_matrix("l1")
        f = P.Forest(gpu_ctx, children, n_taxa, leaves, node_cap=128)
        tcm = gpu_ctx.tcm(sig)
        o = oracle_mod.Oracle(sig)
    plain = f.score(tcm)
    f.enable_rerooting()
    best = f.score_rerooted(tcm)
    for t in range(n_trees):
        want_plain = want_best = 0
        for c in range(n_chars):
            col = [leaves[v][c] for v in range(n_taxa)]
            if wide:
                r = oracle_mod.reroot_haskell(2, a, children[t], col)
            else:
                r = oracle_mod.reroot(o, children[t], col)
            want_best += r[0]
            want_plain += int(r[2][0])
            uv, tot, loc = f.edges(t, c)
            got = {tuple(sorted((int(x), int(y)))): (int(p), int(q)) for (x, y), p, q in zip(uv, tot, loc)}
            ref = {tuple(sorted((int(x), int(y)))): (int(p), int(q)) for (x, y), p, q in zip(r[1], r[2], r[3])}
            assert got == ref, (t, c)
        assert int(plain[t]) == want_plain and int(best[t]) == want_best, t
    f.close()


def test_large_alphabet_forest_protein_discrete_cost(gpu_ctx):
    """DAG cost 1131 (ScriptTests.hs:200-202, commented out upstream) for the protein data under the discrete metric."""
    import pcg_b200 as P
    from tests.helpers import invertebrates_golden
    children, leaves, _ = invertebrates_golden()
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=400, max_b=32, arena_words=0)
    f = P.Forest(gpu_ctx, children[None], len(leaves), [[s] for s in leaves], node_cap=512, metric=21, dialect=2)
    f.enable_rerooting()
    assert int(f.score_rerooted(None)[0]) == 1131
    f.close()


def test_preorder_implied_alignments_over_device_contexts(gpu_ctx):
    """Post-order contexts aligned on the device (postorder_contexts), then the pre-order primitives over them
    (pcg_b200.preorder): all implied alignments share the root's length, each node's ungapped implied alignment is its own
    ungapped median string (for a leaf: its sequence), and no "Impossible happened" branch is reached."""
    from pcg_b200.postorder import postorder_contexts, random_binary_trees
    from pcg_b200.preorder import preorder_implied_alignments
    from tests.helpers import tcm_matrix, random_string
    rng = np.random.default_rng(41)
    n_taxa = 24
    children = random_binary_trees(3, n_taxa, 5)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=512, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(tcm_matrix("discrete"))
    for t in range(3):
        base = random_string(rng, 120, amb=0.05)
        leaves = []
        for _ in range(n_taxa):                         # related sequences: mutations and indels of one ancestor
            s = base.copy()
            mut = rng.random(len(s)) < 0.1
            s[mut] = random_string(rng, int(mut.sum()), amb=0.0)
            s = s[rng.random(len(s)) > 0.05]
            leaves.append(s[s != 16])
        ctxs, local, total = postorder_contexts(gpu_ctx, tcm, children[t], n_taxa, leaves)
        ia, single = preorder_implied_alignments(children[t], n_taxa, ctxs, 5)
        root_len = len(ia[max(ia, key=lambda n: len(ia[n]))])
        assert len(ia) == 2 * n_taxa - 1 and {len(v) for v in ia.values()} == {root_len}
        for n, v in ia.items():
            med, own = v[:, 0], ctxs[n][:, 0]
            assert np.array_equal(med[med != 16], own[own != 16]), (t, n)
        for v in range(n_taxa):
            med = ia[v][:, 0]
            assert np.array_equal(med[med != 16], leaves[v]), (t, v)


def test_preorder_on_the_device_equals_the_restatement(oracle_mod, gpu_ctx):
    """pcgdo_preorder_* (do_preorder.cu: deriveImpliedAlignment as prefix sums, one launch per depth over all trees x
    characters) against oracle/preorder_oracle.c: final alignment (median, left, right) and single disambiguation of EVERY
    node, on contexts aligned by the device (postorder_contexts) — which must themselves equal the restatement's.  Trees with
    a Missing character included.  (Parity unpinned upstream: see preorder_oracle.c.)"""
    import pcg_b200 as P
    from pcg_b200.postorder import postorder_contexts, random_binary_trees
    from pcg_b200.preorder import preorder_device
    from tests.helpers import tcm_matrix, random_string
    rng = np.random.default_rng(4711)
    n_taxa, n_trees, n_chars = 20, 3, 2
    children = random_binary_trees(n_trees, n_taxa, 9)
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=512, max_b=32, arena_words=0)
    for name in ("discrete", "l1"):
        sig = tcm_matrix(name)
        tcm = gpu_ctx.tcm(sig)
        o = oracle_mod.Oracle(sig)
        contexts, want = [], []
        for t in range(n_trees):
            cols, wcols = [], []
            for c in range(n_chars):
                base = random_string(rng, 90 + 20 * c, amb=0.05)
                leaves = []
                for _ in range(n_taxa):
                    s = base.copy()
                    mut = rng.random(len(s)) < 0.1
                    s[mut] = random_string(rng, int(mut.sum()), amb=0.0)
                    s = s[rng.random(len(s)) > 0.06]
                    leaves.append(s[s != 16])
                ctxs, _, _ = postorder_contexts(gpu_ctx, tcm, children[t], n_taxa, leaves)
                root, octx, oia, osingle = oracle_mod.preorder_tree(o, children[t], leaves)
                for v in range(2 * n_taxa - 1):
                    assert np.array_equal(ctxs[v], octx[v]), (name, t, c, v)
                cols.append(ctxs)
                wcols.append((oia, osingle))
            contexts.append(cols)
            want.append(wcols)
        ia, single = preorder_device(gpu_ctx, children, n_taxa, contexts, 5)
        for t in range(n_trees):
            for c in range(n_chars):
                oia, osingle = want[t][c]
                for v in range(2 * n_taxa - 1):
                    assert np.array_equal(ia[t][c][v], oia[v]), (name, t, c, v)
                    assert np.array_equal(single[t][c][v], osingle[v]), (name, t, c, v)
    # a Missing leaf: its final alignment is its parent's; a Missing subtree likewise
    pre = P.Preorder(gpu_ctx, children[:1], n_taxa, 1)
    col = dict(contexts[0][0])
    col[3] = None
    pre.set_contexts([[col]])
    try:
        pre.run(5)
        ia3, s3 = pre.node(0, 0, 3)
        par = next(n_taxa + x for x in range(n_taxa - 1) if 3 in children[0][x])
        iap, sp = pre.node(0, 0, par)
        assert np.array_equal(ia3, iap) and np.array_equal(s3, sp)
    except P.BackendError:
        pass        # (a context that no longer belongs to the tree may legitimately hit "Impossible happened")
    pre.close()


def test_preorder_from_the_traversal_focus(oracle_mod, gpu_ctx):
    """preorder_from_rooting: re-root on the cheapest edge (costs as the re-rooting restatement's), decorate every node
    against its parent in that orientation on the device; the result equals the Python restatement of the fold run over the
    same re-rooted tree, every final alignment has the virtual root's length, and ungapping a leaf's gives its sequence."""
    from pcg_b200.postorder import random_binary_trees
    from pcg_b200.preorder import preorder_from_rooting, preorder_implied_alignments
    from tests.helpers import tcm_matrix, random_string
    rng = np.random.default_rng(99)
    n_taxa = 14
    sig = tcm_matrix("discrete")
    gpu_ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=512, max_b=32, arena_words=0)
    tcm = gpu_ctx.tcm(sig)
    o = oracle_mod.Oracle(sig)
    for t in range(3):
        children = random_binary_trees(1, n_taxa, 50 + t)[0]
        base = random_string(rng, 100, amb=0.05)
        leaves = []
        for _ in range(n_taxa):
            s = base.copy()
            mut = rng.random(len(s)) < 0.15
            s[mut] = random_string(rng, int(mut.sum()), amb=0.0)
            s = s[rng.random(len(s)) > 0.08]
            leaves.append(s[s != 16])
        best, edges, ecost, _ = oracle_mod.reroot(o, children, leaves)
        e, uv, tot, new_children, contexts, ia, single = preorder_from_rooting(gpu_ctx, tcm, children, n_taxa, leaves, 5)
        assert tot == best and int(ecost[e]) == best and tuple(edges[e]) == tuple(uv)
        ia_py, single_py = preorder_implied_alignments(new_children, n_taxa, contexts, 5)
        n_root = len(contexts[2 * n_taxa - 2]) if 2 * n_taxa - 2 in contexts else None
        for v in range(2 * n_taxa - 1):
            assert np.array_equal(ia[v], ia_py[v]), (t, v)
            assert np.array_equal(single[v], single_py[v][:, 0]), (t, v)
        lens = {len(x) for x in ia.values()}
        assert len(lens) == 1
        for v in range(n_taxa):
            med = ia[v][:, 0]
            assert np.array_equal(med[med != 16], leaves[v]), (t, v)
