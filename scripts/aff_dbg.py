import sys, numpy as np
sys.path.insert(0, '.')
import pcg_b200 as P
from oracle import pyoracle as po
from tests.helpers import random_pair, random_string, tcm_matrix
ctx = P.Context(0)
sig = tcm_matrix("discrete"); go=3; variant=1
o = po.Oracle(sig, go)
ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=700, max_b=32, arena_words=0)
tcm = ctx.tcm(sig, gap_open=go, affine_variant=variant)
rng = np.random.default_rng(303)
pairs = [random_pair(rng, 300, amb=0.2) for _ in range(150)]
for n1, n2 in [(1, 1), (2, 1), (1, 30), (39, 39), (40, 40), (41, 41), (42, 40), (80, 41), (200, 35), (600, 90), (130, 131), (640, 600)]:
    pairs.append((random_string(rng, n1, amb=0.2), random_string(rng, n2, amb=0.2)))
pp = P.PackedPairs.from_lists([p[0] for p in pairs], [p[1] for p in pairs])
r1 = P.align2d_batch(ctx, tcm, pp, want_cells=True)
r2 = P.align2d_batch(ctx, tcm, pp, want_cells=True)
print("deterministic:", np.array_equal(r1.cost, r2.cost), np.array_equal(r1.gapped, r2.gapped))
nb=0
for k,(s1,s2) in enumerate(pairs):
    oc = o.align2d_affine(s1, s2, variant)
    gc = r1.pair(pp, k)
    okc = gc[0]==oc[0]; oks = all(np.array_equal(x,y) for x,y in zip(gc[1:], oc[1:4]))
    if not (okc and oks):
        nb+=1
        if nb<15:
            L,S=max(len(s1),len(s2)),min(len(s1),len(s2)); wq=max(40,L-S+8)+41
            print(k, len(s1), len(s2), "wq", wq, "cost", gc[0], oc[0], "strings_ok", oks, "len", len(gc[1]), len(oc[1]))
print("bad", nb, "of", len(pairs))
