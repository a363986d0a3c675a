#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r04o_slab_tests.log 2>&1; echo "slab tests rc=$?"
tail -5 gpurun_out/r04o_slab_tests.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also"
run() { name=$1; shift; $B "$@" > gpurun_out/r04o_bench_$name.json 2> gpurun_out/r04o_bench_$name.err; echo "$name rc=$?"; }
run default
run sym0 --option sym 0
run bb16 --option bucket_bits 16
run bb17 --option bucket_bits 17
run bb14 --option bucket_bits 14
for f in gpurun_out/r04o_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["config"]["rows"], d["config"]["rows_sha256"][:12])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r04o_all_tests.log 2>&1; echo "all tests rc=$?"
tail -5 gpurun_out/r04o_all_tests.log
