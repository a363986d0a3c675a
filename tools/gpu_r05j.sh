#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
timeout 600 python bench.py --no-cpu-baseline --no-also --steps 20 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r05j/bench.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("ms", round(d["ms_per_step"],3), "e2e", round(e["ms_per_step"],3), {k:round(v,3) for k,v in e["stage_ms"].items()}, "parsed", round(e["parsed_sequences"]["ms_per_step"],3))
PY
