#!/usr/bin/env python3
"""Pin the C oracle's rows on the FULL-SIZE BASELINE panels (config 2 and 3: 20+20 genomes x 5 Mbp) as count + sha256.

    python tests/golden/make_fullsize.py [c2|c3] [threads] [key parts]   ->   tests/golden/fullsize.json (merged)

The oracle (oracle/krisp_oracle.c) holds every k-mer as ASCII: config 2 needs ~25 GB of host memory, config 3 (124-mers) ~55 GB —
generate them where that fits; tests/test_gpu_fullsize.py compares the CUDA path with these digests (sorted rows joined by '\\n').
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SHAPES = {"c2": (20, 20, 5_000_000, 25, 1, 2), "c3": (20, 20, 5_000_000, 32, 60, 32)}


def main():
    from krisp_b200.panel import make_panel
    from oracle import oracle
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    n_in, n_out, glen, L, D, R = SHAPES[which]
    gs = make_panel(n_in, n_out, glen)
    recs = [[r.tobytes() for r in g.records] for g in gs]
    nparts = int(sys.argv[3]) if len(sys.argv) > 3 else 1          # key-hash parts (memory: 1 / nparts of the k-mers at a time)
    t0 = time.time()
    rows, total = [], 0
    for part in range(nparts):
        oracle.set_key_partition(part, nparts)
        r, counts = oracle.search_records(recs, [g.name for g in gs], {g.name for g in gs if g.is_ingroup}, True, L, D, R, nthreads=threads)
        rows += r
        total += int(sum(int(c) for c in counts))
        print(f"part {part + 1}/{nparts}: {len(r)} rows, {time.time() - t0:.0f} s", file=sys.stderr, flush=True)
    oracle.set_key_partition(0, 1)
    rows.sort()
    counts = [total]
    dt = time.time() - t0
    out = os.path.join(HERE, "fullsize.json")
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[which] = {"panel": f"{n_in}+{n_out} genomes x {glen} bp, make_panel defaults (seed 1000)", "L": L, "D": D, "R": R, "rows": len(rows),
                   "rows_sha256": hashlib.sha256("\n".join(rows).encode()).hexdigest(), "records": int(sum(int(c) for c in counts)),
                   "oracle_seconds": round(dt, 1), "oracle_threads": threads, "oracle_key_parts": nparts}
    with open(out, "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
    print(which, data[which])


if __name__ == "__main__":
    main()
