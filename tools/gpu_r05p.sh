#!/bin/bash
# final check of the committed tree: GPU suite, smoke, headline bench line
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05p; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --no-also > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r05p/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["traffic_note"][:60], d["whole_step"]["frac_of_peak"], d["e2e"]["ms_per_step"], d["stage_ms"])
PY
