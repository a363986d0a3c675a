#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r04r_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r04r_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-also --steps 10 > gpurun_out/r04r_bench_n1.json 2> gpurun_out/r04r_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r04r_bench_n1.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("ms", d["ms_per_step"], "e2e", e["ms_per_step"], e["value"], "parsed", e["parsed_sequences"])
print(e["stage_ms"], e["host_ms"])
PY
