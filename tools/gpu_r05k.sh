#!/bin/bash
# round-2 final evidence on one GPU: tests, smoke, bench (ours + reference arm), launch list, ncu --set full of K1+L0 / L1 / K3
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r05k; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also"
timeout 600 $CMD > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"kb_hash_warp|kb_extract_part|kb_part_kernel" -s 3 -c 3 -o $O/prof $CMD > $O/ncu_full.log 2>&1
echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kb_fa_count|kb_fa_pack|kb_rank" -s 2 -c 4 -o $O/prof_small $CMD > $O/ncu_small.log 2>&1
echo "ncu small rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r05k/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks") if k in d}); print(d["roofline"]); print(d["whole_step"])
e=d["e2e"]; print({k:e[k] for k in e if k not in ("stage_ms","host_ms")}); print(e.get("stage_ms"))
print([ (a["config"][:12], a["value"], a["e2e"]["value"]) for a in d.get("also",[])]); print(d.get("cpu_baseline"))
PY
ls -la $O
