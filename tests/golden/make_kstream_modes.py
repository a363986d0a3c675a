#!/usr/bin/env python3
"""Golden vectors for the kstream option combinations beyond krisp_fasta's own (tests/golden/kstream_modes.json), from the
UNMODIFIED reference (build container only, needs /root/reference):

    python tests/golden/make_kstream_modes.py

Every case = the reference's ``kstream`` command line on one of the committed FASTA fixtures; stored: the flags, the line count,
the sha256 of the whole output and its first / last lines (the tables themselves stay out of the repo).
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_runner  # noqa: E402

CASES = [
    # (name, file, flags)
    ("forward_plain_28", "c1/ingroup0.fasta.gz", ["-k", 28, "--map-softmask", "--disallow", "Nn", "--sort"]),
    ("canonical_plain_28", "c1/ingroup0.fasta.gz", ["-k", 28, "--canonicals", "--map-softmask", "--disallow", "Nn", "--sort"]),
    ("canonical_split_allow_28", "c1/ingroup0.fasta.gz",
     ["-k", 28, "--canonicals", "--map-softmask", "--allow", "ACGT", "--split", 25, -2, "--sort", "--sort-cols", 0, 2]),
    ("forward_split_omit_15", "panels/edge/alpha.fasta",
     ["-k", 15, "--omit-softmask", "--disallow", "Nn", "--split", 9, -5, "--sort", "--sort-cols", 0, 2]),
    ("both_plain_allow_12", "c1/outgroup1.fasta.gz", ["-k", 12, "--complements", "--map-softmask", "--allow", "ACGT", "--sort"]),
    ("canonical_plain_omit_allow_15", "panels/edge/beta.fna.gz", ["-k", 15, "--canonicals", "--omit-softmask", "--allow", "ACGTacgt", "--sort"]),
    ("forward_split_allow_21", "panels/p_spacer_3x3/ingroup1.fasta.gz",
     ["-k", 21, "--map-softmask", "--allow", "ACGT", "--split", 10, -10, "--sort", "--sort-cols", 0, 2]),
]


def main():
    assert ref_runner.available(), "needs /root/reference"
    out = []
    for name, rel, flags in CASES:
        text = ref_runner.kstream([os.path.join(HERE, rel)] + list(flags))
        lines = text.splitlines()
        out.append({"name": name, "file": rel, "flags": [str(x) for x in flags], "count": len(lines),
                    "sha256": hashlib.sha256(text.encode()).hexdigest(), "head": lines[:3], "tail": lines[-3:]})
        print(f"  {name}: {len(lines)} lines", file=sys.stderr)
    with open(os.path.join(HERE, "kstream_modes.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_kstream_modes.py", "reference": "grunwaldlab/krisp 0.1.6 (unmodified, /root/reference)",
                   "cases": out}, fh, indent=1)


if __name__ == "__main__":
    main()
