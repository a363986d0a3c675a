#!/bin/bash
# BASELINE config 5 on one GPU: genome-count sweep 10 -> 200 genomes x 5 Mbp, 25/1/2
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05s; mkdir -p $O
: > $O/sweep_n1.jsonl
for G in 5 10 25 50 100; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also --genomes $G $G >> $O/sweep_n1.jsonl 2> $O/sweep_$G.err; echo "G=$G rc=$?"
done
python - <<'PY'
import json
for ln in open("gpurun_out/r05s/sweep_n1.jsonl"):
    d=json.loads(ln); e=d["e2e"]
    print(d["config"]["workload"][:60], "Gbp", d["config"]["total_bases"]/1e9, "ms", round(d["ms_per_step"],2), "Gbp/s", round(d["value"],1), "e2e", round(e["value"],1), "rows", d["config"]["rows"], "passes", d["config"]["radix_passes"], {k:round(v,2) for k,v in d["stage_ms"].items()})
PY
