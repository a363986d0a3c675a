#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05i; mkdir -p $O
timeout 600 python bench.py --no-cpu-baseline --no-also --steps 20 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r05i/bench.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("ms", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, "e2e", round(e["ms_per_step"],3), {k:round(v,3) for k,v in e["stage_ms"].items()}, e["host_ms"], "parsed", round(e["parsed_sequences"]["ms_per_step"],3), "launches", d["gpu_launches"], d["whole_step"]["frac_of_peak"], d["whole_step"]["frac_of_peak_at_round1_bytes"])
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "list rc=$?"
