"""Shared generators for the tests (TCMs of the reference's bench data, random strings)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tcm_matrix(name: str, a: int = 5) -> np.ndarray:
    """Symbol-change matrices used by the reference's own tests / bench data (gap = last symbol)."""
    idx = np.arange(a)
    if name == "discrete":          # bench/strings/dna/dna-discrete.tcm
        return (1 - np.eye(a, dtype=np.uint32)).astype(np.uint32)
    if name == "l1":                # bench/strings/dna/dna-L1-norm.tcm
        return np.abs(np.subtract.outer(idx, idx)).astype(np.uint32)
    if name == "1:2":               # substitution 1, indel 2 (Pairwise/Test.hs metrics)
        m = (1 - np.eye(a, dtype=np.uint32)).astype(np.uint32)
        m[-1, :-1] = 2
        m[:-1, -1] = 2
        return m
    if name == "2:1":
        m = 2 * (1 - np.eye(a, dtype=np.uint32)).astype(np.uint32)
        m[-1, :-1] = 1
        m[:-1, -1] = 1
        return m
    if name == "weird":             # symmetric, non-metric, distinct entries
        assert a == 5
        return np.array([[0, 3, 1, 4, 2], [3, 0, 2, 1, 5], [1, 2, 0, 6, 2], [4, 1, 6, 0, 3], [2, 5, 2, 3, 0]],
                        np.uint32)
    raise KeyError(name)


def random_string(rng, n, a=5, amb=0.15):
    """Elements over an a-symbol alphabet: mostly single non-gap symbols, `amb` fraction arbitrary
    non-empty ambiguity sets (may contain the gap bit), like the reference's generators."""
    x = (1 << rng.integers(0, a - 1, size=n)).astype(np.uint32)
    m = rng.random(n) < amb
    return np.where(m, rng.integers(1, 1 << a, size=n), x).astype(np.uint8)


def random_pair(rng, max_len=400, a=5, amb=0.15):
    n1 = int(rng.integers(1, max_len))
    if rng.random() < 0.6:
        n2 = max(1, n1 + int(rng.integers(-30, 30)))
    else:
        n2 = int(rng.integers(1, max_len))
    s1 = random_string(rng, n1, a, amb)
    if rng.random() < 0.5 and n2 <= n1:
        s2 = s1[:n2].copy()
        mut = rng.random(n2) < 0.1
        s2 = np.where(mut, random_string(rng, n2, a, amb), s2).astype(np.uint8)
    else:
        s2 = random_string(rng, n2, a, amb)
    return s1, s2


def bench_strings_golden():
    """tests/golden/bench_strings_dna.json (made by make_bench_strings_golden.py from the reference's
    bench/strings/dna inputs and the compiled reference aligner).  Returns (encoded ungapped strings,
    {tcm file: (sigma, records [i, j, cost, out_len, crc_gapped, crc_slot1, crc_slot2])})."""
    import json
    from pcg_b200 import formats
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_strings_dna.json")))
    codes = formats.iupac_dna_codes()
    enc = []
    for s in g["text"]:
        e = formats.encode_fasta(s, codes)
        enc.append(e[e != 16])
    return enc, {k: (np.array(g["tcms"][k], np.uint32), g["pairs"][k]) for k in g["pairs"]}


def crc_u8(a):
    import zlib
    return zlib.crc32(np.asarray(a, np.uint8).tobytes())


def chel_golden_tree():
    """The reference's chel golden tree in the forest's numbering: (golden dict, children[1, 16, 2], per-taxon
    ungapped leaf strings, golden-id -> node-id map)."""
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "chel_golden.json")))
    nodes = g["nodes"]
    leaf_ids = [n["id"] for n in nodes if not n["children"]]
    int_ids = [n["id"] for n in nodes if n["children"]]
    n_taxa = len(leaf_ids)
    to_mine = {gid: k for k, gid in enumerate(leaf_ids)}
    to_mine.update({gid: n_taxa + k for k, gid in enumerate(int_ids)})
    children = np.zeros((1, n_taxa - 1, 2), np.int32)
    for k, gid in enumerate(int_ids):
        l, r = sorted(nodes[gid]["children"])
        children[0, k] = (to_mine[l], to_mine[r])
    leaves = []
    for gid in leaf_ids:
        c = np.array(nodes[gid]["context"], np.uint8)
        leaves.append(c[c[:, 0] != 16][:, 0].copy())
    return g, children, leaves, to_mine


def invertebrates_golden():
    """tests/golden/invertebrates_protein.json -> (children[21, 2], leaves (uint32 strings, a = 21), {name: (sigma, cost)})."""
    import json
    from pcg_b200 import formats
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "invertebrates_protein.json")))
    seqs = dict(formats.read_fasta(g["fasta"]))
    names = sorted(seqs)
    codes = formats.iupac_amino_codes()
    leaves = [formats.encode_fasta(seqs[n], codes, np.uint32) for n in names]
    children = formats.read_newick_binary(g["tree"], names)
    cases = {}
    for name, text in g["tcm"].items():
        alphabet, sigma = formats.read_tcm(text)
        assert alphabet == formats.AMINO
        cases[name] = (sigma, g["dag_cost"][name])
    return children, leaves, cases


def missing_values_golden():
    """tests/golden/missing_values_dna.json -> (children, ungapped leaves (uint8, a = 5), sigma, asserted DAG cost)."""
    import json
    from pcg_b200 import formats
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "missing_values_dna.json")))
    seqs = dict(formats.read_fasta(g["fasta"]))
    names = sorted(seqs)
    codes = formats.iupac_dna_codes()
    leaves = []
    for n in names:
        e = formats.encode_fasta(seqs[n], codes, np.uint8)
        leaves.append(e[e != 16])
    _, sigma = formats.read_tcm(g["tcm"])
    return formats.read_newick_binary(g["tree"], names), leaves, sigma, g["dag_cost"]


def custom_alphabet_golden(ds):
    """tests/golden/custom_alphabet_scripts.json -> (a, children, leaves ((n, W) uint32 each), {matrix name: (metric kind,
    factored sigma, weight, asserted DAG cost)}) for one of "slashes" / "large-mix" / "huge-mix".  metric kind as the oracle
    numbers them: 0 discrete, 1 explicit (Sankoff), 2 L1 norm — chosen by pcg_b200.formats.diagnose_tcm like PCG does."""
    import base64
    import json
    import zlib
    from pcg_b200 import formats
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "custom_alphabet_scripts.json")))["sets"][ds]
    alphabet = g["alphabet"]
    a = len(alphabet) + 1
    seqs = dict(formats.read_fastc(g["fasta"]))
    names = sorted(seqs)
    leaves = [formats.encode_fastc_wide(seqs[n], alphabet) for n in names]
    children = formats.read_newick_binary(g["tree"], names)
    cases = {}
    for name, c in g["cases"].items():
        sig = np.frombuffer(zlib.decompress(base64.b64decode(c["sigma_u16_zlib_b64"])), "<u2").reshape(a, a)
        weight, f, kind = formats.diagnose_tcm(sig)
        cases[name] = ({"nonadditive": 0, "explicit": 1, "additive": 2}[kind], f, weight, c["dag_cost"])   # (c["asserted"]: upstream state)
    return a, children, leaves, cases


def multiblock_missing_golden():
    """tests/golden/multiblock_missing_dna.json -> (children, [(leaves with None for a taxon missing from the block, sigma)
    per block], asserted DAG cost).  Leaves are ungapped uint8 strings, a = 5."""
    import json
    import re
    from pcg_b200 import formats
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "multiblock_missing_dna.json")))
    names = sorted(set(re.findall(r"[A-Za-z]+", g["tree"])))
    codes = formats.iupac_dna_codes()
    blocks = []
    for b in g["blocks"]:
        seqs = dict(formats.read_fasta(b["fasta"]))
        leaves = []
        for n in names:
            if n in seqs:
                e = formats.encode_fasta(seqs[n], codes, np.uint8)
                leaves.append(e[e != 16])
            else:
                leaves.append(None)
        blocks.append((leaves, formats.read_tcm(b["tcm"])[1]))
    return formats.read_newick_binary(g["tree"], names), blocks, g["dag_cost"]
