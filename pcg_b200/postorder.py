"""Host-side helpers for post-order tree scoring (tree generation for synthetic workloads, level-by-level
scoring through the Haskell-level wrapper for full contexts).

Reference: the post-order driver `postorderSequence'` (Bio/Graph/PhylogeneticDAG/Postorder.hs:98-158) and
`directOptimizationPostorderPairwise` (DO/Internal.hs:127-140).  The fast path is `pcg_b200.Forest`
(device-resident, cost + ungapped medians); `postorder_contexts` here keeps the full (median,left,right)
alignment contexts with gaps, one GPU launch per level, for reporting and for the golden test.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .pairwise import foreign_pairwise_do_batch


def random_binary_trees(n_trees: int, n_taxa: int, seed: int) -> np.ndarray:
    """Random-join (Yule-like) rooted binary trees: repeatedly join two random live lineages.
    Returns children[n_trees, n_taxa-1, 2] (node ids: leaves 0..n_taxa-1, internal x -> n_taxa + x)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.zeros((n_trees, n_taxa - 1, 2), np.int32)
    for t in range(n_trees):
        live = list(range(n_taxa))
        for x in range(n_taxa - 1):
            i, j = rng.choice(len(live), size=2, replace=False)
            a, b = live[i], live[j]
            out[t, x] = (min(a, b), max(a, b))
            for k in sorted((i, j), reverse=True):
                live.pop(k)
            live.append(n_taxa + x)
    return out


def postorder_contexts(ctx, tcm, children: np.ndarray, n_taxa: int, leaves: Sequence[np.ndarray]
                       ) -> Tuple[Dict[int, np.ndarray], Dict[int, int], Dict[int, int]]:
    """One tree, one character: full alignment contexts of every node (leaf context = (v,v,v),
    DO/Internal.hs:109-121), local and total costs.  Levels are batched through foreignPairwiseDO."""
    ctxs: Dict[int, Optional[np.ndarray]] = {v: np.stack([leaves[v]] * 3, axis=1).astype(np.uint8)
                                             for v in range(n_taxa)}
    local: Dict[int, int] = {v: 0 for v in range(n_taxa)}
    total: Dict[int, int] = {v: 0 for v in range(n_taxa)}
    pending = {n_taxa + x: (int(children[x, 0]), int(children[x, 1])) for x in range(n_taxa - 1)}
    while pending:
        ready = [n for n, (l, r) in pending.items() if l in ctxs and r in ctxs]
        res = foreign_pairwise_do_batch(ctx, tcm, [ctxs[pending[n][0]] for n in ready],
                                        [ctxs[pending[n][1]] for n in ready])
        for n, (cost, c) in zip(ready, res):
            l, r = pending.pop(n)
            ctxs[n] = c
            local[n] = cost
            total[n] = cost + total[l] + total[r]
    return ctxs, local, total
