#!/bin/bash
# N GPUs: sharded parity tests, then the bench with exchange variants (--diag)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out/r05d_n$N; mkdir -p $O
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --diag > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep '"diag"' $O/bench.err > $O/diag.jsonl
python - $O <<'PY'
import json,sys
O=sys.argv[1]
d=json.loads(open(O+"/bench.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "ms", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d.get("parity",{}).get("ok"), "nvlink", d.get("nvlink"))
print({k:round(v,3) for k,v in d["stage_ms"].items()})
for ln in open(O+"/diag.jsonl"):
    r=json.loads(ln); print(r.get("diag"), r.get("ms"), r.get("exchange_ms"), r.get("error"), r.get("timeline"))
    print("   ", {k:v for k,v in (r.get("stage_ms") or {}).items() if "group " not in k})
PY
