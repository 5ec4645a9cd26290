"""GPU: kernel time of the S2 Ukkonen dialect (a = 21, discrete) on unrelated / related / mixed pairs for several
skip-ahead factors (PCGDO_SKIP_FACTOR, 0 = plain doubling).  Costs must not change."""
import os
import subprocess
import sys

import numpy as np

if len(sys.argv) > 1:
    sys.path.insert(0, ".")
    import pcg_b200 as P
    from pcg_b200.pairwise import metric_do_batch
    rng = np.random.default_rng(1)
    a, n, L = 21, 20000, 400
    def rnd(k): return (1 << rng.integers(0, 20, size=k)).astype(np.uint32)
    sets = {}
    sets["unrelated"] = [(rnd(L), rnd(L)) for _ in range(n)]
    rel = []
    for _ in range(n):
        s = rnd(L); t = s.copy()
        mut = rng.random(L) < float(sys.argv[1])
        t[mut] = rnd(int(mut.sum()))
        t = t[rng.random(L) > 0.03]
        rel.append((s, t))
    sets["related"] = rel
    ctx = P.Context(0)
    ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=512, max_b=32)
    for name, pairs in sets.items():
        f, s2 = [p[0] for p in pairs], [p[1] for p in pairs]
        metric_do_batch(ctx, a, 2, f[:2000], s2[:2000])
        r = metric_do_batch(ctx, a, 2, f, s2)
        print(f"  {name:10s} kernel {ctx.last_kernel_ms()[0]:8.2f} ms  cost sum {int(r.cost.sum())}  final offsets {np.bincount(np.log2(np.maximum(r.final_offset,1)).astype(int)).tolist()}")
else:
    for mut in ("0.02", "0.1", "0.3"):
        for fac in ("0", "16", "32", "64", "1000000"):
            print(f"mutation rate {mut}, skip factor {fac}", flush=True)
            subprocess.run([sys.executable, __file__, mut], env={**os.environ, "PCGDO_SKIP_FACTOR": fac} if fac != "0" else {**os.environ, "PCGDO_NO_SKIP_AHEAD": "1"})
