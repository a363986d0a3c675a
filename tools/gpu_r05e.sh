#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05e; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also"
timeout 600 $CMD > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "list rc=$?"
