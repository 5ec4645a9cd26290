#!/usr/bin/env python
"""Generates tests/golden/multiblock_missing_dna.json: inputs and asserted DAG cost of the reference's script test
    scriptCheckCost 45  "dynamic/multi-block/missing/missing.pcg"        (test/TestSuite/ScriptTests.hs:184-186)
— two character blocks over five DNA taxa (IUPAC codes), block A under shared-data/tcms/dna-g2t2.tcm without taxon Epsilon,
block B under dna-g4t4.tcm without taxon Delta: a taxon absent from a block holds a Missing dynamic character there
(handleMissingCharacter, DO/Pairwise/Internal.hs:247-259).  The asserted number is the sum over the blocks of the
re-rooted cost (16 + 29).  Run in the build container (needs /root/reference).
"""
import json
import os
import re

SRC = "/root/reference/test/data-sets"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multiblock_missing_dna.json")


def main():
    src = open("/root/reference/test/TestSuite/ScriptTests.hs").read()
    m = re.search(r"scriptCheckCost\s+(\d+)\s+\"dynamic/multi-block/missing/missing.pcg\"", src)
    d = os.path.join(SRC, "dynamic", "multi-block", "missing")
    t = os.path.join(SRC, "shared-data", "tcms")
    out = {"source": "test/data-sets/dynamic/multi-block/missing/{missing-A.fasta,missing-B.fasta,topology.tree}, "
                     "shared-data/tcms/dna-g2t2.tcm, dna-g4t4.tcm; expected cost from test/TestSuite/ScriptTests.hs:184-186",
           "tree": open(os.path.join(d, "topology.tree")).read().strip(),
           "blocks": [{"fasta": open(os.path.join(d, "missing-A.fasta")).read(), "tcm": open(os.path.join(t, "dna-g2t2.tcm")).read()},
                      {"fasta": open(os.path.join(d, "missing-B.fasta")).read(), "tcm": open(os.path.join(t, "dna-g4t4.tcm")).read()}],
           "dag_cost": int(m.group(1))}
    json.dump(out, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, out["dag_cost"])


if __name__ == "__main__":
    main()
