#!/usr/bin/env python
"""Generates tests/golden/missing_values_dna.json: inputs and asserted DAG cost of the reference's script test
    scriptCheckCost 28  "dynamic/single-block/missing/missing-values.pcg"        (test/TestSuite/ScriptTests.hs:191-193)
— five DNA taxa with IUPAC codes and '?' elements, TCM shared-data/tcms/dna-g4t4.tcm (transition 1, transversion 4,
indel 16), tree topology.tree; alphabet 5 <= 8, so PCG aligns with the foreign C aligner (align2d).
Run in the build container (needs /root/reference):  python tests/golden/make_missing_values_golden.py
"""
import json
import os
import re

SRC = "/root/reference/test/data-sets"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "missing_values_dna.json")


def main():
    src = open("/root/reference/test/TestSuite/ScriptTests.hs").read()
    m = re.search(r"scriptCheckCost\s+(\d+)\s+\"dynamic/single-block/missing/missing-values.pcg\"", src)
    d = os.path.join(SRC, "dynamic", "single-block", "missing")
    out = {"source": "test/data-sets/dynamic/single-block/missing/{missing-values.fasta,topology.tree}, "
                     "shared-data/tcms/dna-g4t4.tcm; expected cost from test/TestSuite/ScriptTests.hs",
           "fasta": open(os.path.join(d, "missing-values.fasta")).read(),
           "tree": open(os.path.join(d, "topology.tree")).read().strip(),
           "tcm": open(os.path.join(SRC, "shared-data", "tcms", "dna-g4t4.tcm")).read(),
           "dag_cost": int(m.group(1))}
    json.dump(out, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, out["dag_cost"])


if __name__ == "__main__":
    main()
