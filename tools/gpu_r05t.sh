#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05t; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -8 $O/tests.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r05t/bench_n1.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), e["host_ms"], "parsed", round(e["parsed_sequences"]["ms_per_step"],3), d["whole_step"]["frac_of_peak_at_round1_bytes"])
print([ (a["config"][:12], round(a["ms_per_step"],2), round(a["value"],1), round(a["e2e"]["ms_per_step"],2), a["rows"], a["rows_sha256"][:8]) for a in d.get("also",[])])
PY
tail -5 $O/bench_n1.err
