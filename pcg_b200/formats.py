"""Input formats either side of the DO path: FASTA / FASTC sequence files, TCM text files, IUPAC codes.

Reference (lib = /root/reference/lib):
* FASTA  — lib/file-parsers/src/File/Format/Fasta/Parser.hs: ``>`` + taxon name line, then sequence lines;
           one character per symbol, whitespace ignored.
* FASTC  — lib/file-parsers/src/File/Format/Fastc/Parser.hs: same header; symbols are whitespace-separated
           tokens, ``[a b c]`` is an ambiguity group.
* TCM    — lib/file-parsers/src/File/Format/TransitionCostMatrix/Parser.hs: first line = the alphabet symbols
           (gap is implicit and LAST), then an (n+1) x (n+1) matrix, row/column n = gap.
* IUPAC  — lib/alphabet/src/Data/Alphabet/IUPAC.hs:62-124 (DNA: lower case = the upper-case set plus gap,
           ``?`` = everything including gap; amino acids: B = DN, Z = EQ, X = all 20, ``?`` = all + gap).

Elements are encoded as the DO path wants them: bit i = symbol i of the alphabet, gap = highest bit
(``1 << (a - 1)``, DYN/Internal.hs:253-257).  Pure host-side text handling: no alignment work happens here.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

DNA = ["A", "C", "G", "T"]
AMINO = list("ACDEFGHIKLMNPQRSTVWY")

_DNA_IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT",
              "M": "AC", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}
_AMINO_IUPAC = {**{c: c for c in AMINO}, "B": "DN", "Z": "EQ", "X": "".join(AMINO)}


def iupac_dna_codes() -> Dict[str, int]:
    """char -> 5-bit element (A=1, C=2, G=4, T=8, gap=16)."""
    bit = {s: 1 << i for i, s in enumerate(DNA)}
    gap = 1 << len(DNA)
    out = {}
    for ch, syms in _DNA_IUPAC.items():
        v = sum(bit[s] for s in syms)
        out[ch] = v
        if ch not in "N":
            out[ch.lower()] = v | gap
    out["-"] = gap
    out["?"] = (gap << 1) - 1
    return out


def iupac_amino_codes() -> Dict[str, int]:
    bit = {s: 1 << i for i, s in enumerate(AMINO)}
    gap = 1 << len(AMINO)
    out = {ch: sum(bit[s] for s in syms) for ch, syms in _AMINO_IUPAC.items()}
    out["-"] = gap
    out["?"] = (gap << 1) - 1
    return out


def read_fasta(text: str) -> List[Tuple[str, str]]:
    """[(taxon name, sequence characters with whitespace removed)]."""
    out: List[Tuple[str, str]] = []
    name, buf = None, []
    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            if name is not None:
                out.append((name, "".join(buf)))
            name, buf = line[1:].strip(), []
        else:
            buf.append("".join(line.split()))
    if name is not None:
        out.append((name, "".join(buf)))
    return out


def encode_fasta(seq: str, codes: Dict[str, int], dtype=np.uint8) -> np.ndarray:
    return np.fromiter((codes[c] for c in seq), dtype=dtype, count=len(seq))


def read_fastc(text: str) -> List[Tuple[str, List[List[str]]]]:
    """[(taxon name, [ambiguity group = list of symbols])]."""
    out: List[Tuple[str, List[List[str]]]] = []
    name, toks = None, []

    def flush():
        if name is None:
            return
        groups, cur = [], None
        for t in toks:
            while t:
                if t[0] == "[":
                    cur, t = [], t[1:]
                elif t[-1] == "]" and cur is not None:
                    if t[:-1]:
                        cur.append(t[:-1])
                    groups.append(cur)
                    cur, t = None, ""
                else:
                    (cur if cur is not None else groups).append(t if cur is not None else [t])
                    t = ""
        out.append((name, groups))

    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            flush()
            name, toks = line[1:].strip(), []
        else:
            toks.extend(line.split())
    flush()
    return out


def encode_fastc(groups: Sequence[Sequence[str]], alphabet: Sequence[str], dtype=np.uint32) -> np.ndarray:
    """alphabet WITHOUT the gap; '-' encodes as the gap bit, '?' as everything."""
    bit = {s: 1 << i for i, s in enumerate(alphabet)}
    gap = 1 << len(alphabet)
    bit["-"] = gap
    full = (gap << 1) - 1
    return np.fromiter((full if g == ["?"] else sum(bit[s] for s in g) for g in groups), dtype=dtype, count=len(groups))


def encode_fastc_wide(groups: Sequence[Sequence[str]], alphabet: Sequence[str]) -> np.ndarray:
    """As ``encode_fastc`` for alphabets of any size: (n, W) uint32, W = ceil((len(alphabet) + 1) / 32); bit i of an
    element = bit i % 32 of word i // 32 (the layout of the 32-bit-word forest, ``pcgdo_forest_create32`` with a > 32)."""
    a = len(alphabet) + 1
    W = (a + 31) // 32
    idx = {s: i for i, s in enumerate(alphabet)}
    idx["-"] = a - 1
    out = np.zeros((len(groups), W), np.uint32)
    for k, g in enumerate(groups):
        for i in (range(a) if list(g) == ["?"] else (idx[s] for s in g)):
            out[k, i >> 5] |= np.uint32(1 << (i & 31))
    return out


def diagnose_tcm(sigma: np.ndarray) -> Tuple[int, np.ndarray, str]:
    """``diagnoseTcm`` as far as the DO dispatch needs it (lib/core/tcm/src/Data/TCM/Internal.hs:506-524, 549-590;
    used by ``dynamicMetadata``, Bio/Metadata/Dynamic/Internal.hs:286-300): (weight, factored matrix, structure) with
    structure ``"additive"`` (sigma(i,j) = |i-j| -> ``firstLinearNormPairwiseLogic``), ``"nonadditive"`` (0 on / 1 off the
    diagonal -> ``discreteMetricPairwiseLogic``) or ``"explicit"`` (memoized Sankoff overlap).  The common factor of all
    entries is divided out first (``factorTCM``); the DO runs on the factored matrix and PCG multiplies the weight back."""
    m = np.asarray(sigma, dtype=np.int64)
    n = m.shape[0]
    factor = int(np.gcd.reduce(m.reshape(-1)))
    f = m // factor if factor > 1 else m
    i, j = np.indices((n, n))
    # |i-j| and 0/1 matrices are symmetric metrics (the 0/1 one ultrametric too), so the guards above them in diagnoseTcm
    # always pass; Additive is tested first, which matters for the 2 x 2 matrix that is both
    if bool((f == np.abs(i - j)).all()):
        kind = "additive"
    elif bool((f == (i != j)).all()):
        kind = "nonadditive"
    else:
        kind = "explicit"
    return max(factor, 1), f.astype(np.uint32), kind


def read_tcm(text: str) -> Tuple[List[str], np.ndarray]:
    """(alphabet without gap, (n+1) x (n+1) uint32 matrix with the gap as last row / column)."""
    lines = [ln.split() for ln in text.splitlines() if ln.strip()]
    alphabet = lines[0]
    n = len(alphabet) + 1
    rows = lines[1:1 + n]
    if len(rows) != n or any(len(r) != n for r in rows):
        raise ValueError(f"TCM: expected a {n} x {n} matrix after the alphabet line")
    return alphabet, np.array([[int(float(x)) for x in r] for r in rows], dtype=np.uint32)


def read_newick_binary(text: str, taxa: Sequence[str]):
    """A rooted binary Newick tree over `taxa` -> children[(n-1), 2] int32 in the forest's numbering (leaf = index in
    `taxa`, internal node x = len(taxa) + x, children listed before parents).  Branch lengths and labels of internal
    nodes are ignored (lib/file-parsers/src/File/Format/Newick)."""
    import re
    idx = {n: i for i, n in enumerate(taxa)}
    n_taxa = len(taxa)
    s = "".join(text.split()).rstrip(";")
    children: List[List[int]] = []

    def skip_annot(pos):
        m = re.match(r"[^,():;]*(:[^,():;]*)?", s[pos:])
        return pos + len(m.group(0))

    def parse(pos):
        if s[pos] == "(":
            kids, pos = [], pos + 1
            while True:
                k, pos = parse(pos)
                kids.append(k)
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError("malformed Newick")
            if len(kids) != 2:
                raise ValueError("not a binary tree")
            children.append(kids)
            return n_taxa + len(children) - 1, skip_annot(pos)
        m = re.match(r"[^,():;]+", s[pos:])
        name = m.group(0)
        return idx[name], skip_annot(pos + len(name)) if s[pos + len(name):pos + len(name) + 1] == ":" else pos + len(name)

    parse(0)
    if len(children) != n_taxa - 1:
        raise ValueError("tree does not span the taxa")
    return np.array(children, dtype=np.int32)
