#!/bin/bash
# tail rework (rank rows, one D2H, lazy group sizes) + K3 with 10 warps per CTA
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/tests.log
run() { tag=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-also --steps 10 "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench $tag rc=$?"
python - $tag <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r05b/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
e=d["e2e"]
print(sys.argv[1], "ms", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, "e2e", round(e["ms_per_step"],3), "parsed", round(e["parsed_sequences"]["ms_per_step"],3), "launches", d["gpu_launches"])
PY
}
run w10
run w8 --option hash_warps 8
run w10s10 --option hash_slots_log2 10
