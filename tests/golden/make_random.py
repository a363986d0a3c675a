#!/usr/bin/env python3
"""More pins for the CPU oracle: small random panels (random group sizes, lengths, L/D/R, soft masking, N runs, duplications) run
through the UNMODIFIED reference (build container only, needs /root/reference):

    python tests/golden/make_random.py

Writes the FASTA files under tests/golden/random/<case>/ and the reference's rows to tests/golden/random.json.  tests/test_oracle.py
checks the C oracle and the Python set model against them (CPU); the GPU suite keeps to the cases it was validated on.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from krisp_b200.panel import make_panel, write_panel  # noqa: E402
from make_golden import run_case  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    rdir = os.path.join(HERE, "random")
    shutil.rmtree(rdir, ignore_errors=True)
    cases = []
    for i in range(14):
        n_in, n_out = int(rng.integers(1, 4)), int(rng.integers(0, 4))
        glen = int(rng.integers(800, 2500))
        L, D, R = int(rng.integers(0, 13)), int(rng.integers(0, 5)), int(rng.integers(0, 9))
        if L + R == 0:
            L = 4
        omit = bool(rng.integers(0, 2))
        flags = {"conserved-left": L, "diagnostic": D, "conserved-right": R}
        name = f"r{i:02d}_{n_in}x{n_out}_{L}_{D}_{R}" + ("_omit" if omit else "")
        genomes = make_panel(n_in, n_out, glen, seed=int(rng.integers(1, 1 << 30)), snp_every=int(rng.integers(40, 200)),
                             noise=float(rng.choice([0.0, 1e-3, 1e-2])), n_runs=int(rng.integers(0, 3)), run_len=int(rng.integers(1, 30)),
                             dup_len=int(rng.integers(0, 200)), soft_frac=float(rng.choice([0.0, 0.05, 0.3])), soft_block=int(rng.integers(5, 80)),
                             n_records=int(rng.integers(1, 4)))
        i_p, o_p = write_panel(genomes, os.path.join(rdir, name), compress=bool(rng.integers(0, 2)))
        cases.append(run_case(name, i_p, o_p, flags, omit_soft=omit))
    with open(os.path.join(HERE, "random.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_random.py", "reference": "grunwaldlab/krisp 0.1.6 (unmodified, /root/reference)", "cases": cases}, fh, indent=1)
    print(f"wrote {len(cases)} cases, {sum(c['n_rows'] for c in cases)} rows in total", file=sys.stderr)


if __name__ == "__main__":
    main()
