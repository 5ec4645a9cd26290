"""GPU check: a forest over a block of characters must give, per tree, the sum of those characters' costs in the full forest."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench as B
import pcg_b200 as P
from pcg_b200.sharding import character_block

n_trees, n_taxa, n_chars, leaf_len = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 256
world = int(sys.argv[4])
children, leaves = B.po_inputs(n_trees, n_taxa, n_chars, leaf_len)
ctx = P.Context(0)
ctx.configure(pairs_per_warp=32, warps_per_cta=32, max_len=B.PO["node_cap"], max_b=32)
tcm = ctx.tcm(B.discrete_tcm())
full = P.Forest(ctx, children, n_taxa, [[leaves[v, c] for c in range(n_chars)] for v in range(n_taxa)], B.PO["node_cap"])
want = full.score(tcm)
full.close()
tot = np.zeros(n_trees, np.int64)
for r in range(world):
    lo, hi = character_block(n_chars, r, world)
    f = P.Forest(ctx, children, n_taxa, [[leaves[v, c] for c in range(lo, hi)] for v in range(n_taxa)], B.PO["node_cap"])
    got = f.score(tcm)
    tot += got
    print("rank", r, "chars", lo, hi, "sum", int(got.sum()), flush=True)
    f.close()
print("full", int(want.sum()), "sharded", int(tot.sum()), "equal", bool((tot == want).all()))
