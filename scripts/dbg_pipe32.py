import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcg_b200 as P
from pcg_b200.pairwise import metric_do_batch
a = 21
rng = np.random.default_rng(a)
def rstr(n):
    sym = rng.integers(0, a - 1, size=n)
    return (1 << (sym % 32)).astype(np.uint32)
pairs = [(rstr(int(rng.integers(1, 120))), rstr(int(rng.integers(1, 120)))) for _ in range(1500)]
f, s = [p[0] for p in pairs], [p[1] for p in pairs]
ctx = P.Context(0)
ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=128, max_b=32, arena_words=0)
os.environ["PCGDO_HOST_CHUNK_PAIRS"] = "1048576"; ctx.reload_env()
plain = metric_do_batch(ctx, a, 2, f, s)
for chunk in ("256", "512", "256"):
    os.environ["PCGDO_HOST_CHUNK_PAIRS"] = chunk; ctx.reload_env()
    piped = metric_do_batch(ctx, a, 2, f, s)
    bad = []
    for k in range(len(pairs)):
        x, y = plain.pair(k), piped.pair(k)
        if x[0] != y[0] or any(not np.array_equal(u, v) for u, v in zip(x[1:], y[1:])):
            bad.append(k)
    print("chunk", chunk, "bad pairs", len(bad), bad[:10], bad[-5:], "n_elems", len(plain.median),
          "first diff elem", int(np.argmax(plain.median != piped.median)) if not np.array_equal(plain.median, piped.median) else None)
