"""Extract the known-answer vectors of the reference's own golden test into a small JSON fixture.

Source (read-only, NOT available on the GPU box):
  /root/reference/test/data-sets/golden-tests/chel-golden-test/{chel_xml.golden,chel_data.golden}
  (compared byte-for-byte by the reference's test/TestSuite/GoldenTests.hs:38-60).

Run once in the build container:  python tests/golden/make_chel_golden.py
Writes tests/golden/chel_golden.json:
  {"alphabet": 5, "tcm": "discrete", "nodes": [{"id", "children", "local_cost", "total_cost", "context": [[m,l,r],...]}]}
"""
import json
import os
import re

SRC = "/root/reference/test/data-sets/golden-tests/chel-golden-test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chel_golden.json")


def main():
    xml = open(os.path.join(SRC, "chel_xml.golden")).read()
    final = xml.split("<Final_graph>", 1)[1]
    chunks = final.split("<Phylogenetic_node>")[1:]
    assert len(chunks) == 33
    topo = {}
    for line in open(os.path.join(SRC, "chel_data.golden")):
        m = re.match(r"^\s+(\d+)\s+\{([\d,]*)\}\s+\{([\d,]*)\}\s*$", line)
        if m:
            topo[int(m.group(1))] = [int(x) for x in m.group(3).split(",") if x]
        if len(topo) == 33:
            break
    nodes = []
    for i, ch in enumerate(chunks):
        local = int(re.search(r"<Local_cost>\s*&(\d+);", ch).group(1))
        total = int(re.search(r"<Character_cost>\s*&(\d+);", ch).group(1))
        ctx = re.search(r"<Alignment_Context>\s*&DC \[(.*?)\];", ch, re.S).group(1)
        triples = [[int(a), int(b), int(c)]
                   for a, b, c in re.findall(r"\(\[5\](\d+),\[5\](\d+),\[5\](\d+)\)", ctx)]
        nodes.append({"id": i, "children": topo[i], "local_cost": local, "total_cost": total,
                      "context": triples})
    json.dump({"alphabet": 5, "tcm": "discrete", "source": "chel_xml.golden <Final_graph>",
               "rerooted_nodes": [0, 30], "nodes": nodes}, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
