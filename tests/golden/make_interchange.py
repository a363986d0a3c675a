#!/usr/bin/env python3
"""Golden interchange files: what the UNMODIFIED reference hands from its search stages to its renderer.

Run in the build container only (needs /root/reference):

    python tests/golden/make_interchange.py

For a subset of the cases of golden.json it runs the reference (oracle/ref_runner.krisp_fasta_interchange, which only observes
the file name passed to render_output) and stores the text of ``filtered.txt`` / ``merged_file.txt`` (krisp_fasta.py:256-283):
lines ``left,mid,right,label(n);label...`` (Amplicon.py:170-206, :298-348), groups in ascending (left, right) order.
Written to tests/golden/interchange.json.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_runner  # noqa: E402

CASES = ["c1_spacer_25_1_2", "c1_conserved30_diag0", "c1_conserved30_amplicon100", "c1_no_outgroup", "p_spacer_3x3", "p_spacer_3x3_omit",
         "p_spacer_4x5", "p_5_2_3", "p_9_4_3", "p_12_3_12", "p_0_2_6", "p_33_2_31", "p_16_0_16", "edge_strand_crlf", "edge_multicopy",
         "edge_label_collision", "edge_multicopy_shared", "p_primer_3x3"]
MAX_BYTES = 150_000


def main():
    assert ref_runner.available(), "needs /root/reference"
    with open(os.path.join(HERE, "golden.json")) as fh:
        golden = json.load(fh)
    by_name = {c["name"]: c for c in golden["cases"]}
    out = {}
    for name in CASES:
        c = by_name[name]
        argv = [os.path.join(HERE, p) for p in c["ingroup"]]
        if c["outgroup"]:
            argv += ["--outgroup"] + [os.path.join(HERE, p) for p in c["outgroup"]]
        for k, v in c["flags"].items():
            argv += [f"--{k}", str(v)]
        if c["omit_soft"]:
            argv.append("--omit-soft")
        argv += ["--cores", "1"]
        stdout, text = ref_runner.krisp_fasta_interchange(argv)
        assert sorted(ref_runner.rows_of(stdout)) == sorted(c.get("rows", ref_runner.rows_of(stdout)))
        if len(text) > MAX_BYTES:
            print(f"  {name}: {len(text)} bytes, skipped", file=sys.stderr)
            continue
        out[name] = text
        print(f"  {name}: {len(text.splitlines())} lines", file=sys.stderr)
    with open(os.path.join(HERE, "interchange.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
