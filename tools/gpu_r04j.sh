#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$1
for CS in 1 3; do
KRISP_COPY_STREAMS=$CS timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2952$CS bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r04j_bench_n${N}_cs$CS.json 2> gpurun_out/r04j_bench_n${N}_cs$CS.err; echo "bench n$N CS=$CS rc=$?"
python - $N $CS <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r04j_bench_n{sys.argv[1]}_cs{sys.argv[2]}.json").read().strip().splitlines()[-1])
sm=d["stage_ms"]
print("N",sys.argv[1],"CS",sys.argv[2],"ms", round(d["ms_per_step"],3), "value", round(d["value"],2), "e2e", round(d["e2e"]["ms_per_step"],3), "K1", round(sm.get("K1 extract + partition 0",0),3), "xchg", round(sm.get("K4 exchange (bulk peer copies, first send to last landed)",0),3), "L1", round(sum(v for k,v in sm.items() if k.startswith("K2 partition 1")),3), "K3", round(sum(v for k,v in sm.items() if k.startswith("K3")),3), d["config"]["rows"])
PY
done
