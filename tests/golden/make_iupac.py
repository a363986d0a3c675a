#!/usr/bin/env python3
"""Golden cases for letters other than ACGTN (edge behaviour E1) from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_iupac.py

A 3+3 x 20 kbp seeded panel (25/1/2) is edited at planted group-SNP sites:
  out_mid   an IUPAC letter in the MIDDLE column of one outgroup genome  -> the reference keeps the k-mer, the letter counts
            as a base of its own (disjoint from every ingroup base): the row is still printed (kstream.py:11-18,715-732)
  flank     an IUPAC letter in a conserved FLANK of one outgroup genome   -> that genome lacks the flank key: the row dies at the
            intersection, with or without the letter being "kept"
  in_mid    an IUPAC letter in the MIDDLE column of one ingroup genome    -> the reference's render worker raises KeyError
            (Amplicon.py:65) and the run still exits 0 with truncated output: undefined behaviour, recorded as observed
Writes tests/golden/panels/iupac_*/ and tests/golden/iupac.json.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_runner  # noqa: E402
from krisp_b200.panel import make_panel, write_panel  # noqa: E402

L, D, R = 25, 1, 2


def rel(p):
    return os.path.relpath(p, HERE)


def run(ins, outs):
    argv = list(ins) + ["--outgroup"] + list(outs) + ["--conserved-left", L, "--diagnostic", D, "--conserved-right", R, "--cores", 1]
    stdout, stderr = ref_runner.krisp_fasta(argv)
    return ref_runner.rows_of(stdout), stderr


def main():
    base = make_panel(3, 3, 20_000, n_records=1, n_runs=0, soft_frac=0.0, dup_len=0, noise=0.0)
    d0 = os.path.join(HERE, "panels", "iupac_base")
    ins, outs = write_panel(base, d0)
    rows0, _ = run(ins, outs)
    assert len(rows0) >= 20, len(rows0)
    # a planted site in the middle of the genome whose forward-strand row exists: window = [p - 25, p + 3)
    seq0 = base[0].records[0].tobytes().decode()
    sites = [p for p in range(500, 20_000, 1000) if any(r.startswith(seq0[p - L:p] + ",") for r in rows0)]
    p1, p2, p3 = sites[3], sites[7], sites[11]
    cases = []

    def edited(name, gi, pos, letter):
        gs = make_panel(3, 3, 20_000, n_records=1, n_runs=0, soft_frac=0.0, dup_len=0, noise=0.0)
        gs[gi].records[0][pos] = ord(letter)
        d = os.path.join(HERE, "panels", "iupac_" + name)
        return write_panel(gs, d)

    for name, gi, pos, letter in (("out_mid", 3, p1, "R"), ("out_mid_lower", 4, p1, "y"), ("flank", 3, p2 + 1, "Y"), ("in_mid", 0, p3, "K")):
        i2, o2 = edited(name, gi, pos, letter)
        rows, stderr = run(i2, o2)
        cases.append({"name": name, "ingroup": [rel(p) for p in i2], "outgroup": [rel(p) for p in o2], "L": L, "D": D, "R": R,
                      "genome": gi, "position": pos, "letter": letter, "reference_rows": rows,
                      "reference_keyerror": "KeyError" in stderr,
                      "rows_without_letter": rows0,
                      "rows_lost_by_reference": sorted(set(rows0) - set(rows)), "rows_gained_by_reference": sorted(set(rows) - set(rows0))})
        print(name, len(rows0), "->", len(rows), "KeyError" in stderr, file=sys.stderr)
    with open(os.path.join(HERE, "iupac.json"), "w") as fh:
        json.dump({"base": {"ingroup": [rel(p) for p in ins], "outgroup": [rel(p) for p in outs], "rows": rows0}, "cases": cases}, fh, indent=1)


if __name__ == "__main__":
    main()
