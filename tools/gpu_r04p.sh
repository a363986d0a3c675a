#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$1
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r04p_bench_n${N}_$tag.json 2> gpurun_out/r04p_bench_n${N}_$tag.err; echo "bench n$N $tag rc=$?"
tail -2 gpurun_out/r04p_bench_n${N}_$tag.err
python - $N $tag <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r04p_bench_n{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
sm=d["stage_ms"]
print("N",sys.argv[1],sys.argv[2],"ms", round(d["ms_per_step"],3), "value", round(d["value"],2), "e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"],2), {k:round(v,3) for k,v in sm.items() if "group " not in k}, d["config"]["rows"], "parity", d["parity"]["ok"])
print("nvlink", {k:v for k,v in d.get("nvlink",{}).items() if k!="how"})
PY
}
run copy1 KRISP_SLAB_EXCHANGE=copy KRISP_COPY_STREAMS=1
run copy7 KRISP_SLAB_EXCHANGE=copy KRISP_COPY_STREAMS=7
run a2a_g4 KRISP_SLAB_EXCHANGE=a2a KRISP_SLAB_GROUPS=4
