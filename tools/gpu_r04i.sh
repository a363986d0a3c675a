#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_worker.py > gpurun_out/r04i_worker2.log 2>&1; echo "worker rc=$?"
tail -40 gpurun_out/r04i_worker2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r04i_bench_n2.json 2> gpurun_out/r04i_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r04i_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r04i_bench_n2.json").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["stage_ms"], d["config"]["rows"])
PY
for G in 1 2 4; do
KRISP_SLAB_GROUPS=$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r04i_bench_n2_g$G.json 2> gpurun_out/r04i_bench_n2_g$G.err; echo "bench n2 G=$G rc=$?"
python - $G <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r04i_bench_n2_g{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("G",sys.argv[1],"ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["stage_ms"], d["config"]["rows"])
PY
done
