#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It (1) copies the reference's five test FASTA fixtures (test_data/krisp_fasta,
data only, no source) to tests/golden/c1/, (2) writes small seeded synthetic
panels to tests/golden/panels/<case>/, (3) runs the unmodified reference
(oracle/ref_runner.py) on every case and stores the CSV rows (canonically
sorted), the --out_align text where requested, and per-file stage-A/B k-mer
tables (count + sha256 + gz of the reference's own ``*.{k}mers`` content via
``kstream --sort``) in tests/golden/golden.json.
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_runner  # noqa: E402
from krisp_b200.panel import make_panel, write_panel, Genome  # noqa: E402

REF_DATA = "/root/reference/test_data/krisp_fasta"


def rel(p):
    return os.path.relpath(p, HERE)


def run_case(name, ingroup, outgroup, flags, out_align=False, dot=False, omit_soft=False, cores=1):
    argv = list(ingroup)
    if outgroup:
        argv += ["--outgroup"] + list(outgroup)
    for k, v in flags.items():
        argv += [f"--{k}", str(v)]
    if omit_soft:
        argv.append("--omit-soft")
    argv += ["--cores", str(cores)]
    align_text = None
    with tempfile.TemporaryDirectory() as td:
        if out_align:
            ap = os.path.join(td, "align.txt")
            argv += ["--out_align", ap]
            if dot:
                argv.append("--dot-alignment")
        stdout, _ = ref_runner.krisp_fasta(argv)
        if out_align:
            with open(ap) as fh:
                align_text = fh.read()
    rows = ref_runner.rows_of(stdout)
    case = {"name": name, "ingroup": [rel(p) for p in ingroup], "outgroup": [rel(p) for p in outgroup],
            "flags": flags, "omit_soft": omit_soft, "dot": dot, "n_rows": len(rows),
            "rows_sha256": hashlib.sha256("\n".join(rows).encode()).hexdigest()}
    if len(rows) <= 3000:          # big row sets are pinned by count + digest only
        case["rows"] = rows
    if align_text is not None:
        case["out_align"] = align_text
    print(f"  {name}: {len(rows)} rows", file=sys.stderr)
    return case


def kstream_table(path, L, D, R, omit_soft=False):
    """Reference stage A+B table of one file (what extractSortedKmers writes)."""
    k = L + D + R
    argv = [path, "-k", k, "--complements", "--disallow", "Nn",
            "--omit-softmask" if omit_soft else "--map-softmask",
            "--split", L, -R, "--sort", "--sort-cols", 0, 2]
    out = ref_runner.kstream(argv)
    lines = out.splitlines()
    return {"file": rel(path), "L": L, "D": D, "R": R, "omit_soft": omit_soft, "count": len(lines),
            "sha256": hashlib.sha256(out.encode()).hexdigest(),
            "head": lines[:5], "tail": lines[-5:]}, out


def write_fasta(path, records, names=None, width=60, crlf=False):
    nl = "\r\n" if crlf else "\n"
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt", newline="") as fh:
        for j, r in enumerate(records):
            fh.write(f">{names[j] if names else 'rec%d' % j}{nl}")
            for s in range(0, len(r), width):
                fh.write(r[s:s + width] + nl)
            if len(r) == 0:
                fh.write(nl)


def main():
    assert ref_runner.available(), "needs /root/reference"
    cases, tables = [], []

    # ---- C1: the reference's own fixture + README worked examples ---------------------------
    c1 = os.path.join(HERE, "c1")
    os.makedirs(c1, exist_ok=True)
    for f in sorted(os.listdir(REF_DATA)):
        shutil.copyfile(os.path.join(REF_DATA, f), os.path.join(c1, f))
        os.chmod(os.path.join(c1, f), 0o644)
    ins = [os.path.join(c1, f"ingroup{i}.fasta.gz") for i in range(2)]
    outs = [os.path.join(c1, f"outgroup{i}.fasta.gz") for i in range(3)]
    sp = {"conserved-left": 25, "diagnostic": 1, "conserved-right": 2}
    cases.append(run_case("c1_spacer_25_1_2", ins, outs, sp, out_align=True))
    cases.append(run_case("c1_spacer_25_1_2_dot", ins, outs, sp, out_align=True, dot=True))
    cases.append(run_case("c1_conserved30_diag0", ins + outs, [], {"conserved": 30, "diagnostic": 0}))
    cases.append(run_case("c1_conserved30_amplicon100", ins, outs, {"conserved": 30, "amplicon": 100}, out_align=True))
    cases.append(run_case("c1_primer_32_60_32", ins, outs, {"conserved-left": 32, "diagnostic": 60, "conserved-right": 32}))
    cases.append(run_case("c1_spacer_cores3", ins, outs, sp, cores=3))
    cases.append(run_case("c1_no_outgroup", ins, [], sp))
    cases.append(run_case("c1_single_file", ins[:1], [], {"conserved-left": 12, "diagnostic": 2, "conserved-right": 12}))
    t, text = kstream_table(ins[0], 25, 1, 2)
    tables.append(t)
    with gzip.open(os.path.join(HERE, "c1_ingroup0_28mers.txt.gz"), "wt") as fh:
        fh.write(text)
    t, _ = kstream_table(outs[1], 32, 60, 32)
    tables.append(t)

    # ---- seeded synthetic panels (same generator as bench.py, scaled down) -----------------
    pdir = os.path.join(HERE, "panels")
    shutil.rmtree(pdir, ignore_errors=True)
    specs = [
        # name, n_in, n_out, len, seed, flags, omit
        ("p_spacer_3x3", 3, 3, 6000, 11, sp, False),
        ("p_spacer_3x3_omit", 3, 3, 6000, 11, sp, True),
        ("p_primer_3x3", 3, 3, 6000, 11, {"conserved-left": 32, "diagnostic": 60, "conserved-right": 32}, False),
        ("p_spacer_4x5", 4, 5, 4000, 12, sp, False),
        ("p_10_1_2", 2, 2, 3000, 13, {"conserved-left": 10, "diagnostic": 1, "conserved-right": 2}, False),
        ("p_6_1_2", 2, 3, 3000, 14, {"conserved-left": 6, "diagnostic": 1, "conserved-right": 2}, False),
        ("p_5_2_3", 2, 2, 3000, 15, {"conserved-left": 5, "diagnostic": 2, "conserved-right": 3}, False),
        ("p_4_3_4", 3, 1, 3000, 16, {"conserved-left": 4, "diagnostic": 3, "conserved-right": 4}, True),
        ("p_9_4_3", 1, 4, 3000, 17, {"conserved-left": 9, "diagnostic": 4, "conserved-right": 3}, False),
        ("p_12_3_12", 2, 2, 4000, 18, {"conserved-left": 12, "diagnostic": 3, "conserved-right": 12}, False),
        ("p_8_0_8", 2, 2, 3000, 19, {"conserved": 8, "diagnostic": 0}, False),
        ("p_0_2_6", 2, 2, 3000, 20, {"conserved-left": 0, "diagnostic": 2, "conserved-right": 6}, False),
        ("p_7_1_0_quirk", 2, 2, 3000, 21, {"conserved-left": 7, "diagnostic": 1, "conserved-right": 0}, False),
        ("p_30_40_30", 2, 3, 5000, 22, {"conserved": 30, "amplicon": 100}, False),
        ("p_20_30_20_noout", 3, 0, 4000, 23, {"conserved": 20, "diagnostic": 30}, False),
        ("p_40_70_40", 2, 2, 5000, 24, {"conserved": 40, "diagnostic": 70}, False),
        ("p_16_0_16", 3, 2, 4000, 25, {"conserved": 16, "diagnostic": 0}, False),
        ("p_33_2_31", 2, 2, 4000, 26, {"conserved-left": 33, "diagnostic": 2, "conserved-right": 31}, False),
    ]
    for name, n_in, n_out, glen, seed, flags, omit in specs:
        d = os.path.join(pdir, name)
        genomes = make_panel(n_in, n_out, glen, seed=seed, snp_every=200, noise=2e-3, n_runs=2, run_len=20,
                             dup_len=300, soft_frac=0.05, soft_block=60, n_records=3)
        i_p, o_p = write_panel(genomes, d, compress=True)
        cases.append(run_case(name, i_p, o_p, flags, omit_soft=omit, out_align=(name in ("p_spacer_3x3", "p_5_2_3"))))
    t, _ = kstream_table(os.path.join(pdir, "p_spacer_3x3", "ingroup1.fasta.gz"), 25, 1, 2)
    tables.append(t)
    t, _ = kstream_table(os.path.join(pdir, "p_spacer_3x3_omit", "outgroup0.fasta.gz"), 25, 1, 2, omit_soft=True)
    tables.append(t)
    t, _ = kstream_table(os.path.join(pdir, "p_primer_3x3", "ingroup0.fasta.gz"), 32, 60, 32)
    tables.append(t)

    # ---- hand-made edge cases -------------------------------------------------------------------
    edir = os.path.join(pdir, "edge")
    os.makedirs(edir, exist_ok=True)
    rng = np.random.default_rng(77)

    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))

    core = rnd(400)
    pal = "ACGTTGCAAGCTTGCAACGT" + "GAATTC" * 3           # palindromic stretches
    a = core[:200] + "A" + core[201:] + pal
    b = core[:200] + "C" + core[201:] + pal
    c = core[:200] + "G" + core[201:] + pal
    rc = {"A": "T", "C": "G", "G": "C", "T": "A"}
    c_rev = "".join(rc[x] for x in reversed(c))            # same genome, other strand
    e1 = os.path.join(edir, "alpha.fasta")
    e2 = os.path.join(edir, "beta.fna.gz")
    e3 = os.path.join(edir, "gamma.1.fa")
    e4 = os.path.join(edir, "delta.fasta")
    write_fasta(e1, [a[:250], a[250:], "ACGT", ""], crlf=True)                 # CRLF, short + empty records
    write_fasta(e2, [b.lower()[:100] + b[100:], "NNNN" + rnd(50)])
    write_fasta(e3, [c_rev])
    write_fasta(e4, [c[:300] + "n" + c[301:], rnd(120)])
    cases.append(run_case("edge_strand_crlf", [e1, e2], [e3, e4], {"conserved-left": 9, "diagnostic": 1, "conserved-right": 5}, out_align=True))
    cases.append(run_case("edge_strand_crlf_omit", [e1, e2], [e3, e4], {"conserved-left": 9, "diagnostic": 1, "conserved-right": 5}, omit_soft=True))
    # label collision: an outgroup file whose simplename equals an ingroup label (membership is by label)
    e5 = os.path.join(edir, "alpha.v2.fasta")
    write_fasta(e5, [c])
    cases.append(run_case("edge_label_collision", [e1], [e5, e3], {"conserved-left": 9, "diagnostic": 1, "conserved-right": 5}))
    # multi-copy: the variable site twice in the ingroup genome with different bases
    e6 = os.path.join(edir, "multi.fasta")
    write_fasta(e6, [a, b])
    cases.append(run_case("edge_multicopy", [e6], [e3], {"conserved-left": 11, "diagnostic": 1, "conserved-right": 4}, out_align=True))
    cases.append(run_case("edge_multicopy_shared", [e6], [e4, e2], {"conserved-left": 11, "diagnostic": 1, "conserved-right": 4}))
    # genome shorter than k in one file => no rows
    e7 = os.path.join(edir, "tiny.fasta")
    write_fasta(e7, ["ACGTACGTAC"])
    cases.append(run_case("edge_too_short", [e1], [e7], {"conserved-left": 9, "diagnostic": 1, "conserved-right": 5}))
    t, _ = kstream_table(e1, 9, 1, 5)
    tables.append(t)
    t, _ = kstream_table(e2, 9, 1, 5, omit_soft=True)
    tables.append(t)

    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "grunwaldlab/krisp 0.1.6 (unmodified, import stubs from oracle/stubs)",
                   "cases": cases, "tables": tables}, fh, indent=1)
    print(f"wrote {len(cases)} cases, {len(tables)} tables", file=sys.stderr)


if __name__ == "__main__":
    main()
