#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py -x -q -m gpu -k "host_buffers or fasta or batch or golden" > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline --no-also --steps 10 > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench $tag rc=$?"
python - $tag <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r05g/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
e=d["e2e"]
print(sys.argv[1], "ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), {k:round(v,3) for k,v in e["stage_ms"].items()}, "parsed", round(e["parsed_sequences"]["ms_per_step"],3))
PY
}
run sched X=1
run uniform KRISP_UNIFORM_BATCHES=1
run sched_fa1 KRISP_FA_STREAMS=1
