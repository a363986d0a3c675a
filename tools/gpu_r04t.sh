#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --no-cpu-baseline --no-also --steps 10 > gpurun_out/r04t_bench_$tag.json 2> gpurun_out/r04t_bench_$tag.err; echo "bench $tag rc=$?"
python - $tag <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r04t_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
e=d["e2e"]
print(sys.argv[1], "ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), "parsed", round(e["parsed_sequences"]["ms_per_step"],3), {k:round(v,3) for k,v in e["stage_ms"].items()})
PY
}
run fa1 KRISP_FA_STREAMS=1
run fa2 KRISP_FA_STREAMS=2
run fa4 KRISP_FA_STREAMS=4
run fa8 KRISP_FA_STREAMS=8
run fa8c32 KRISP_FA_STREAMS=8 CUDA_DEVICE_MAX_CONNECTIONS=32
run fa4c32 KRISP_FA_STREAMS=4 CUDA_DEVICE_MAX_CONNECTIONS=32
