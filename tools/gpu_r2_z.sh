#!/bin/bash
# Round-2 call Z (last of the round): validation + bench lines at HEAD after the warp-uniform slot claim, active-lane probe of every stage.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_z.log; }
rm -f $O/steps_z.log
timeout 300 python -m pytest tests -m gpu -q --tb=short > $O/pytest_z.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_z.log 2>&1; step smoke $?
timeout 200 python bench.py --steps 20 --warmup 5 > $O/z_bench_k20.json 2> $O/bench_z_k20.err; step bench_k20 $?
B2ENV_LIB=$PWD/variants/libb2env_lanes.so timeout 120 python tools/stage_profile.py 300,1000 > $O/lanes_z.log 2>&1; step lanes $?
timeout 120 compute-sanitizer --tool memcheck --print-limit 4 python tools/sanitize_case.py > $O/sanitize_z.log 2>&1; step memcheck $?
timeout 300 python bench.py > $O/z_bench_full.json 2> $O/bench_z_full.err; step bench_full $?
timeout 200 python bench.py --workload pandagrasp --steps 200 --warmup 10 > $O/z_bench_pandagrasp.json 2> $O/bench_z_grasp.err; step bench_grasp $?
timeout 200 python bench.py --workload pandareach --steps 200 --warmup 10 > $O/z_bench_pandareach.json 2> $O/bench_z_reach.err; step bench_reach $?
echo done >> $O/steps_z.log
tail -4 $O/pytest_z.log | cut -c1-200; cat $O/smoke_z.log; cat $O/steps_z.log; tail -3 $O/sanitize_z.log; grep "active lanes" $O/lanes_z.log | cut -c1-200
for f in k20 full pandagrasp pandareach; do python - <<PY
import json
try:
    d=json.loads(open("$O/z_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.3f M"%(d["value"]/1e6), "e2e %.3f M"%(d["e2e"]["value"]/1e6), "cpu", (d.get("cpu_baseline") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"), d.get("config",{}).get("kernel_ms_by_replica"))
except Exception as e:
    print("$f failed", e)
PY
done
