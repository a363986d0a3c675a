#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r04l_bench_n${N}.json 2> gpurun_out/r04l_bench_n${N}.err; echo "bench n$N rc=$?"
tail -5 gpurun_out/r04l_bench_n${N}.err
python - $N <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r04l_bench_n{sys.argv[1]}.json").read().strip().splitlines()[-1])
sm=d["stage_ms"]
print("N",sys.argv[1],"ms", round(d["ms_per_step"],3), "value", round(d["value"],2), "e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"],2), {k:round(v,3) for k,v in sm.items() if "group " not in k}, d["config"]["rows"])
print("parity", d.get("parity"))
print("nvlink", d.get("nvlink"))
PY
