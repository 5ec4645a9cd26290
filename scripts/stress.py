"""GPU stress: the parity checks of tests/ with fresh random seeds, for a wall-clock budget (seconds, argv[1]).
Prints the seed of any mismatch and exits 1.  Test infrastructure (uses the oracle); not part of the product."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import pcg_b200 as P
from oracle import pyoracle as O
from pcg_b200.pairwise import metric_do_batch
from pcg_b200.postorder import random_binary_trees
from tests.helpers import random_pair, random_string, tcm_matrix
from tests.test_gpu_metric import _pairs, _rand_str
from tests.test_gpu_wide import _wide_pairs, _sigma

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
t_end = time.time() + budget
ctx = P.Context(0)
seed0 = int(time.time()) % 1000000
n_checked = {"linear": 0, "explicit": 0, "wide": 0, "forest": 0}
it = 0
while time.time() < t_end:
    seed = seed0 + it
    it += 1
    rng = np.random.default_rng(seed)
    try:
        # dense C-linear aligner, compact + plain transfer
        name = ["discrete", "l1", "1:2", "2:1", "weird"][seed % 5]
        sig = tcm_matrix(name)
        o = O.Oracle(sig)
        ctx.configure(pairs_per_warp=int(rng.choice([2, 8, 32])), warps_per_cta=32, max_len=700, max_b=32, arena_words=0)
        tcm = ctx.tcm(sig)
        pairs = [random_pair(rng, int(rng.choice([60, 250, 650]))) for _ in range(1200)]
        pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
        # every other round through the pipelined host path with packed outputs (tiny chunks: 6 chunks, ring of 3 slots)
        compact = bool(seed & 1)
        os.environ["PCGDO_HOST_CHUNK_PAIRS"] = "256" if compact else "1048576"
        ctx.reload_env()
        r = P.align2d_batch(ctx, tcm, pp, want_cells=True, compact=compact, pinned=bool(seed & 2))
        if compact:                          # rounds whose outputs really came back packed (not the plain fallback)
            n_checked["packed_rounds"] = n_checked.get("packed_rounds", 0) + int((r.out_off != pp.out_off).any())
        for k, (s1, s2) in enumerate(pairs):
            oc = o.align2d(s1, s2)
            gc = r.pair(pp, k)
            assert gc[0] == oc[0] and all(np.array_equal(x, y) for x, y in zip(gc[1:], oc[1:4])) and int(r.cells[k]) == oc[4], ("linear", k)
        n_checked["linear"] += len(pairs)
        # Haskell dialects over the three overlap kinds, a <= 32
        a = int(rng.choice([9, 12, 21, 32]))
        kind = ["discrete", "explicit", "l1"][seed % 3]
        sg = _sigma(a, rng) if kind == "explicit" else None
        dialect = int(rng.integers(0, 4))
        ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=300, max_b=32, arena_words=0)
        memo = ctx.memo(sg, kind, a=a)
        pairs = _pairs(rng, a, 150, 220)
        r = metric_do_batch(ctx, memo, dialect, [p[0] for p in pairs], [p[1] for p in pairs])
        for k, (s1, s2) in enumerate(pairs):
            oc = O.haskell_do(dialect, a, s1, s2, sigma=sg, metric_kind={"discrete": 0, "explicit": 1, "l1": 2}[kind])
            gc = r.pair(k)
            assert gc[0] == oc[0] and all(np.array_equal(x, y) for x, y in zip(gc[1:], oc[1:4])) and int(r.final_offset[k]) == oc[5], ("explicit", a, kind, dialect, k)
        n_checked["explicit"] += len(pairs)
        # wide alphabets
        a = int(rng.choice([33, 64, 65, 130, 300]))
        sg = _sigma(a, rng) if kind == "explicit" else None
        dialect = int(rng.integers(0, 3))
        ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=128, max_b=32, arena_words=0)
        memo = ctx.memo(sg, kind, a=a)
        pairs = _wide_pairs(rng, a, 40, 100)
        r = metric_do_batch(ctx, memo, dialect, [p[0] for p in pairs], [p[1] for p in pairs])
        for k, (s1, s2) in enumerate(pairs):
            oc = O.haskell_do_w(dialect, a, s1, s2, sigma=sg, metric_kind={"discrete": 0, "explicit": 1, "l1": 2}[kind])
            gc = r.pair(k)
            assert gc[0] == oc[0] and all(np.array_equal(x, y) for x, y in zip(gc[1:], oc[1:4])) and int(r.final_offset[k]) == oc[5], ("wide", a, kind, dialect, k)
        n_checked["wide"] += len(pairs)
        # forest with Missing / empty leaves, re-rooting
        n_taxa, n_trees, n_chars = int(rng.integers(4, 14)), 4, 4
        children = random_binary_trees(n_trees, n_taxa, seed)
        leaves = [[None if rng.random() < 0.15 else np.zeros(0, np.uint8) if rng.random() < 0.1 else random_string(rng, int(rng.integers(3, 60)))
                   for _ in range(n_chars)] for _ in range(n_taxa)]
        sig = tcm_matrix(["discrete", "l1", "1:2"][seed % 3])
        o = O.Oracle(sig)
        ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=160, max_b=32, arena_words=0)
        f = P.Forest(ctx, children, n_taxa, leaves, node_cap=160)
        f.enable_rerooting()
        best = f.score_rerooted(ctx.tcm(sig))
        for t in range(n_trees):
            want = sum(O.reroot(o, children[t], [leaves[v][c] for v in range(n_taxa)])[0] for c in range(n_chars))
            assert int(best[t]) == want, ("forest", t)
        f.close()
        n_checked["forest"] += n_trees * n_chars
    except AssertionError as ex:
        print("MISMATCH seed", seed, ex.args, flush=True)
        sys.exit(1)
print("stress ok:", it, "rounds,", n_checked, "seed0", seed0)
