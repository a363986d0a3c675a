#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out/r05n_n$N; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/sharded_worker.py > $O/worker.log 2>&1; echo "worker rc=$?"
grep -E "FAIL|three-level|divergent" $O/worker.log | head -20; tail -3 $O/worker.log
