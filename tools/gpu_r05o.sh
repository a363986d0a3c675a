#!/bin/bash
# 8 GPUs: sharded parity worker, then BASELINE config 4 (50+50 x 60 Mbp = 6 Gbp) with the grouped three-level pipeline
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-8}
O=gpurun_out/r05o_n$N; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/sharded_worker.py > $O/worker.log 2>&1; echo "worker rc=$?"
grep -c "^ok" $O/worker.log; grep -E "FAIL|three-level" $O/worker.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --genomes 50 50 --genome-len 7500000 > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
python - $O <<'PY'
import json,sys
O=sys.argv[1]
for f in ("bench_c4.json",):
    try:
        d=json.loads(open(O+"/"+f).read().strip().splitlines()[-1])
        print(f, "N", d["n_gpus"], "ms", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"],1), "parity", d.get("parity",{}).get("ok"), "rows", d["config"]["rows"])
        print("  nvlink", {k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get("nvlink") or {}).items() if k!="how"})
        print("  ", {k:round(v,3) for k,v in d["stage_ms"].items() if "group " not in k})
    except Exception as e: print(f, "ERR", e); print(open(O+"/"+f.replace(".json",".err")).read()[-1500:])
PY
