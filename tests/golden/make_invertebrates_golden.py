#!/usr/bin/env python
"""Generates tests/golden/invertebrates_protein.json: the inputs of two of the reference's own script tests and the DAG
costs those tests assert (test/TestSuite/ScriptTests.hs:199-204):

    scriptCheckCost 1948  "dynamic/single-block/protein/1-2/invertebrates.pcg"
    scriptCheckCost 1242  "dynamic/single-block/protein/2-1/invertebrates.pcg"

Each script reads test/data-sets/shared-data/invertebrates/invertebrates.fasta as ONE dynamic character per taxon over
the custom alphabet of shared-data/tcms/protein-{1-2,2-1}.tcm (20 amino acids + gap: a = 21 > 8, so PCG aligns with
unboxedUkkonenFullSpaceDO over the memoized explicit matrix) and the tree invertebrates.tree; the asserted number is the
cost of the decorated DAG, i.e. the post-order followed by the re-rooting pass.
Run in the build container (needs /root/reference):  python tests/golden/make_invertebrates_golden.py
"""
import json
import os
import re

SRC = "/root/reference/test/data-sets/shared-data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "invertebrates_protein.json")


def main():
    src = open("/root/reference/test/TestSuite/ScriptTests.hs").read()
    want = {}
    for name in ("1-2", "2-1"):
        m = re.search(r"scriptCheckCost\s+(\d+)\s+\"dynamic/single-block/protein/%s/invertebrates.pcg\"" % name, src)
        want[name] = int(m.group(1))
    out = {"source": "test/data-sets/shared-data/invertebrates/{invertebrates.fasta,invertebrates.tree}, "
                     "shared-data/tcms/protein-{1-2,2-1}.tcm; expected costs from test/TestSuite/ScriptTests.hs:199-204",
           "fasta": open(os.path.join(SRC, "invertebrates", "invertebrates.fasta")).read(),
           "tree": open(os.path.join(SRC, "invertebrates", "invertebrates.tree")).read().strip(),
           "tcm": {n: open(os.path.join(SRC, "tcms", f"protein-{n}.tcm")).read() for n in want},
           "dag_cost": want}
    json.dump(out, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes", want)


if __name__ == "__main__":
    main()
