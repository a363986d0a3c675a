#!/bin/bash
# ncu of the kernels of the strand-symmetric level 0 (what >= 4 GPUs run), on one GPU with option sym = 1
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05u; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also --option sym 1"
timeout 300 $CMD > $O/plain.json 2> $O/plain.err; echo "plain rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r05u/plain.json").read().strip().splitlines()[-1])
print("sym=1 on one GPU: ms", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kb_extract_items|kb_part_expand" -s 2 -c 2 -o $O/prof_sym $CMD > $O/ncu.log 2>&1; echo "ncu rc=$?"
