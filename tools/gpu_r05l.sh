#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05l; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -12 $O/tests.log
run() { tag=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-also --steps 10 "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench $tag rc=$?"
python - $tag <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r05l/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "ms", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if "group " not in k}, "e2e", round(d["e2e"]["ms_per_step"],3))
PY
}
run g0
run g2 --option pipe_groups 2
run g4 --option pipe_groups 4
run g8 --option pipe_groups 8
