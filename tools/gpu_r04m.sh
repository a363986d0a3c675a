#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$1
shift
for MODE in "$@"; do
KRISP_SLAB_EXCHANGE=$MODE timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r04m_bench_n${N}_$MODE.json 2> gpurun_out/r04m_bench_n${N}_$MODE.err; echo "bench n$N $MODE rc=$?"
tail -3 gpurun_out/r04m_bench_n${N}_$MODE.err
python - $N $MODE <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r04m_bench_n{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
sm=d["stage_ms"]
print("N",sys.argv[1],sys.argv[2],"ms", round(d["ms_per_step"],3), "value", round(d["value"],2), "e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"],2), {k:round(v,3) for k,v in sm.items() if "group " not in k}, d["config"]["rows"], "parity", d["parity"]["ok"])
print("nvlink", {k:v for k,v in d.get("nvlink",{}).items() if k!="how"})
PY
done
