#!/usr/bin/env python
"""Generates tests/golden/tcm_memo_golden.json: overlap lookups answered by the UNMODIFIED reference tcm-memo
(lib/tcm-memo/ffi/memoized-tcm, compiled by oracle/Makefile into oracle/_ref/libtcmmemo_ref.so) through its
own entry points matrixInit / getCostAndMedian2D, for the reference's protein and DNA TCM files and one
asymmetric random matrix.  Each record: [x, y, median, cost] (ambiguity sets as integers, bit i = symbol i).

Run in the build container (needs /root/reference):  python tests/golden/make_tcm_memo_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from pcg_b200 import formats  # noqa: E402

FILES = ["bench/strings/protein/protein-pref-gap.tcm", "bench/strings/protein/protein-pref-sub.tcm",
         "bench/strings/protein/protein-L1-norm.tcm", "bench/strings/dna/dna-pref-gap.tcm"]


def main():
    pyoracle.build(ref=True)
    rng = np.random.Generator(np.random.PCG64(20260110))
    mats = {}
    for f in FILES:
        _, sg = formats.read_tcm(open(os.path.join("/root/reference", f)).read())
        mats[f] = sg
    mats["random-asymmetric-12"] = rng.integers(0, 9, size=(12, 12)).astype(np.uint32)
    out = {}
    for name, sg in mats.items():
        a = sg.shape[0]
        ref = pyoracle.ReferenceMemo(sg)
        recs = []
        for k in range(600):
            x = int(rng.integers(1, 1 << a))
            y = int(rng.integers(1, 1 << a))
            if k % 3 == 0:
                x = 1 << int(rng.integers(0, a))
            if k % 5 == 0:
                y = 1 << int(rng.integers(0, a))
            if k % 7 == 0:
                y = 1 << (a - 1)
            med, cost = ref.overlap(x, y)
            recs.append([x, y, med, cost])
        out[name] = {"sigma": sg.tolist(), "lookups": recs}
    with open(os.path.join(ROOT, "tests", "golden", "tcm_memo_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print({k: len(v["lookups"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
