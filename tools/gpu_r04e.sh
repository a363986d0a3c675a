#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --option bucket_bits 18 --option hash_slots_log2 8"
timeout 600 $CMD > gpurun_out/r04e_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"kb_hash_warp" -s 1 -c 1 -o gpurun_out/r04e_prof $CMD > gpurun_out/r04e_ncu.log 2>&1
echo "ncu rc=$?"
