#!/usr/bin/env python
"""Generates tests/golden/bench_strings_dna.json from the reference's own benchmark inputs and the
UNMODIFIED reference aligner compiled by oracle/Makefile (oracle/_ref/libpcgdo_ref.so).

BASELINE.json configs[0] ("bench-string-alignment-small"): the 20 strings of
/root/reference/bench/strings/dna/dna-sequences.fasta (a8 ... a4096 unambiguous, r9 ... r4608 IUPAC
ambiguous incl. '-' and '?'), all 190 unordered pairs, under the reference's DNA TCM files.  Each string is
IUPAC-encoded (lib/alphabet/src/Data/Alphabet/IUPAC.hs) and stripped of gap-only elements, as
`measureAndUngapCharacters` does before the FFI call (DO/Pairwise/Internal.hs:364-381); the LONGER string
is passed first (DO/Pairwise/FFI.hsc:261-268).  Recorded per pair: the cost align2d returns and the CRC-32
of the three output strings (gapped median, inputChar1 slot, inputChar2 slot; one byte per element).

Run in the build container (needs /root/reference):  python tests/golden/make_bench_strings_golden.py
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from pcg_b200 import formats  # noqa: E402

SRC = "/root/reference/bench/strings/dna"
TCMS = ["dna-discrete.tcm", "dna-L1-norm.tcm", "dna-pref-gap.tcm", "dna-pref-sub.tcm"]


def crc(a):
    return zlib.crc32(np.asarray(a, np.uint8).tobytes())


def main():
    pyoracle.build(ref=True)
    codes = formats.iupac_dna_codes()
    gap = 16
    seqs = formats.read_fasta(open(os.path.join(SRC, "dna-sequences.fasta")).read())
    names = [n for n, _ in seqs]
    enc = []
    for _, s in seqs:
        e = formats.encode_fasta(s, codes)
        enc.append(e[e != gap])
    out = {"source": "bench/strings/dna/dna-sequences.fasta", "names": names,
           "text": [s for _, s in seqs], "tcms": {}, "pairs": {}}
    for tf in TCMS:
        alphabet, sigma = formats.read_tcm(open(os.path.join(SRC, tf)).read())
        assert alphabet == formats.DNA
        ref = pyoracle.Reference(sigma)
        recs = []
        for i in range(len(enc)):
            for j in range(i + 1, len(enc)):
                a, b = (enc[i], enc[j]) if len(enc[i]) >= len(enc[j]) else (enc[j], enc[i])
                cost, g, s1, s2 = ref.align2d(a, b)
                recs.append([i, j, cost, len(g), crc(g), crc(s1), crc(s2)])
        out["tcms"][tf] = sigma.tolist()
        out["pairs"][tf] = recs
        print(tf, "sum of costs", sum(r[2] for r in recs))
    with open(os.path.join(ROOT, "tests", "golden", "bench_strings_dna.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
