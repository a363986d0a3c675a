#!/usr/bin/env python3
"""Time the UNMODIFIED reference (grunwaldlab/krisp, /root/reference) on the scaled panel of BASELINE.md section 3.

Build container only (the reference does not travel to the GPU box).  20 + 20 genomes x 100 kbp from bench.py's generator,
uncompressed FASTA on local disk, `krisp_fasta ... --cores $(nproc)`, wall clock around the whole process (the reference's only
mode).  Writes one JSON line: Gbp/s, seconds, cores, rows, and whether the rows equal the C oracle's on the same panel.

    python tools/time_reference.py [--genome-len 100000] [--ldr 25 1 2] > profiles/r04_python_reference.json
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-len", type=int, default=100_000)
    ap.add_argument("--ldr", nargs=3, type=int, default=[25, 1, 2])
    ap.add_argument("--genomes", nargs=2, type=int, default=[20, 20])
    a = ap.parse_args()
    from krisp_b200.panel import make_panel, write_panel
    from oracle import oracle, ref_runner
    if not ref_runner.available():
        print(json.dumps({"unavailable": "the reference is not mounted here (build container only)"}))
        return
    L, D, R = a.ldr
    gs = make_panel(a.genomes[0], a.genomes[1], a.genome_len)
    bases = sum(g.n_bases for g in gs)
    with tempfile.TemporaryDirectory() as td:
        ins, outs = write_panel(gs, os.path.join(td, "panel"))
        wd = os.path.join(td, "work")
        os.makedirs(wd)
        argv = list(ins) + ["--outgroup"] + list(outs) + ["--conserved-left", L, "--diagnostic", D, "--conserved-right", R,
                                                           "--cores", os.cpu_count(), "--workdir", wd]
        t0 = time.perf_counter()
        stdout, _ = ref_runner.krisp_fasta(argv)
        dt = time.perf_counter() - t0
    rows = ref_runner.rows_of(stdout)
    recs = [[r.tobytes() for r in g.records] for g in gs]
    t1 = time.perf_counter()
    want, _ = oracle.search_records(recs, [g.name for g in gs], {g.name for g in gs if g.is_ingroup}, True, L, D, R, nthreads=os.cpu_count())
    dt_port = time.perf_counter() - t1
    print(json.dumps({"impl": "unmodified reference (Python, GNU sort, multiprocessing)", "value": bases / dt / 1e9, "unit": "Gbp/s",
                      "seconds": dt, "cores": os.cpu_count(), "kind": "reference",
                      "sample": f"{a.genomes[0]}+{a.genomes[1]} genomes x {a.genome_len} bp ({bases / 1e6:.1f} Mbp), {L}/{D}/{R}, uncompressed FASTA, --cores {os.cpu_count()}",
                      "rows": len(rows), "rows_equal_c_oracle": rows == want,
                      "c_oracle_port_same_panel": {"value": bases / dt_port / 1e9, "seconds": dt_port, "cores": os.cpu_count(), "kind": "port"}}))


if __name__ == "__main__":
    main()
