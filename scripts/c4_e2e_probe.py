"""Times pcgdo_metric_batch_host on config 4 with and without the chunked pipeline (PCGDO_HOST_CHUNK_PAIRS)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
import bench as B
import pcg_b200 as P
from pcg_b200.backend import Batch32
n, L4, A = 500000, 400, 21
ctx = P.Context(0)
ctx.configure(pairs_per_warp=16, warps_per_cta=16, max_len=L4, max_b=32)
strings = B.config4_strings(n)
elems = np.ascontiguousarray(strings.reshape(-1))
k = np.arange(n, dtype=np.uint64)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
off1 = pin(k * np.uint64(2 * L4)); off2 = pin(off1 + np.uint64(L4)); ln = pin(np.full(n, L4, np.uint32)); oo = pin(off1.copy())
cnt = elems.size + 4
h_elems = pin(elems)
h_m, h_a1, h_a2 = (torch.empty(cnt, dtype=torch.int32).pin_memory().numpy() for _ in range(3))
h_cost, h_olen = (torch.empty(n, dtype=torch.int32).pin_memory().numpy() for _ in range(2))
hb = Batch32(n, h_elems.ctypes.data, elems.size, off1.ctypes.data, off2.ctypes.data, ln.ctypes.data, ln.ctypes.data,
             oo.ctypes.data, cnt, h_m.ctypes.data, h_a1.ctypes.data, h_a2.ctypes.data, h_olen.ctypes.data, h_cost.ctypes.data, None, None)
for chunk in ("1048576", "65536", "32768", "131072"):
    os.environ["PCGDO_HOST_CHUNK_PAIRS"] = chunk
    ctx.reload_env()
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.check(ctx.lib.pcgdo_metric_batch_host(ctx.h, A, 2, C.byref(hb)), "host")
        dt = time.perf_counter() - t0
    print("chunk", chunk, "ms", round(dt * 1e3, 1), "checksum", int(h_cost.astype(np.int64).sum()), "launches", ctx.last_kernel_ms())
