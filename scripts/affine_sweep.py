#!/usr/bin/env python
"""Throughput of the affine kernel by band class: pairs of ~5 kb strings with a fixed length difference."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcg_b200 as P
from pcg_b200.backend import Batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = P.Context(0)
ctx.configure(pairs_per_warp=8, warps_per_cta=16, max_len=5400, max_b=32)
tcm = ctx.tcm((1 - np.eye(5, dtype=np.uint32)).astype(np.uint32), gap_open=3, affine_variant=1)
rng = np.random.default_rng(1)
for delta in [int(x) for x in os.environ.get("DELTAS", "0,60,150,300,480").split(",")]:
    L1, L2 = 5000, 5000 - delta
    a = (1 << rng.integers(0, 4, size=(n, L1))).astype(np.uint8)
    b = (1 << rng.integers(0, 4, size=(n, L2))).astype(np.uint8)
    pp = P.PackedPairs.from_lists(list(a), list(b))
    dev = torch.device("cuda", 0)
    t = lambda x: torch.from_numpy(x).to(dev)
    d = [t(pp.elems), t(pp.off1.view(np.int64)), t(pp.off2.view(np.int64)), t(pp.len1.view(np.int32)),
         t(pp.len2.view(np.int32)), t(pp.out_off.view(np.int64))]
    og = torch.empty(pp.out_bytes, dtype=torch.uint8, device=dev); o1 = torch.empty_like(og); o2 = torch.empty_like(og)
    ol = torch.empty(n, dtype=torch.int32, device=dev); co = torch.empty(n, dtype=torch.int32, device=dev)
    ce = torch.empty(n, dtype=torch.int64, device=dev)
    bt = Batch(n, d[0].data_ptr(), pp.elems.nbytes, d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
               d[5].data_ptr(), pp.out_bytes, og.data_ptr(), o1.data_ptr(), o2.data_ptr(), ol.data_ptr(), co.data_ptr(),
               ce.data_ptr())
    for _ in range(2):
        ctx.check(ctx.lib.pcgdo_align2d_affine_batch_device(ctx.h, tcm.h, C.byref(bt), None), "affine")
    torch.cuda.synchronize()
    ms = ctx.last_kernel_ms()[0]
    cells = float(ce.sum().item())
    print(f"delta {delta:4d}  wq_need {max(40, delta + 8) + 41:4d}  {ms:8.2f} ms  {cells / ms / 1e6:8.1f} GCUPS  cost sum {int(co.sum().item())}")
