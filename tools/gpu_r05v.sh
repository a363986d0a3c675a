#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05v; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests.log
timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-also --option sym 1 > $O/sym.json 2> $O/sym.err; echo "sym rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r05v/sym.json").read().strip().splitlines()[-1])
print("sym=1 on one GPU: ms", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["config"]["rows"], d["config"]["rows_sha256"][:8])
PY
