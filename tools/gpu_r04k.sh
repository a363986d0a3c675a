#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r04k_all_tests.log 2>&1; echo "all tests rc=$?"
tail -3 gpurun_out/r04k_all_tests.log
timeout 900 python bench.py > gpurun_out/r04k_bench_n1.json 2> gpurun_out/r04k_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r04k_bench_n1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r04k_bench_ref.json 2> gpurun_out/r04k_bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r04k_bench_n1.json").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
print({k:round(v,3) for k,v in d["stage_ms"].items()})
print("roof", d["roofline"]["kernel"][:30], d["roofline"]["frac"], "whole", d["whole_step"]["frac_of_peak"], d["whole_step"]["frac_of_peak_at_round1_bytes"])
print("fam", d["kernel_families"])
for a in d.get("also", []): print("also", a["config"][:40], a["value"], a["ms_per_step"], a["e2e"], a["rows"], a["frac"])
print("cpu", d.get("cpu_baseline"))
print("launches", d["gpu_launches"], d["clocks"])
PY
