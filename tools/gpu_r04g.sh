#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu > gpurun_out/r04g_slab_tests.log 2>&1; echo "slab tests rc=$?"
tail -5 gpurun_out/r04g_slab_tests.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r04g_all_tests.log 2>&1; echo "all tests rc=$?"
tail -5 gpurun_out/r04g_all_tests.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; $B "$@" > gpurun_out/r04g_bench_$name.json 2> gpurun_out/r04g_bench_$name.err; echo "$name rc=$?"; }
run default
run bb16_s10_cta --option bucket_bits 16 --option hash_slots_log2 10
run bb16_s11_cta --option bucket_bits 16 --option hash_slots_log2 11
run bb17_s9_cta --option bucket_bits 17 --option hash_slots_log2 9 --option hash_shared 1
run bb17_s10_cta --option bucket_bits 17 --option hash_slots_log2 10
run bb15_s11_cta --option bucket_bits 15 --option hash_slots_log2 11
run bb15_s12_cta --option bucket_bits 15 --option hash_slots_log2 12
run bb18_s8_cta --option bucket_bits 18 --option hash_slots_log2 8 --option hash_shared 1
for f in gpurun_out/r04g_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["config"]["rows"], d["config"]["k3_stats"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
