#!/bin/bash
# Round-2 call X: validation + bench lines at HEAD after the serial-motor-rows change.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_x.log; }
rm -f $O/steps_x.log
timeout 600 python -m pytest tests -m gpu -q --tb=short > $O/pytest_x.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_x.log 2>&1; step smoke $?
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1_final_k20.json 2> $O/bench_x_k20.err; step bench_k20 $?
timeout 400 python bench.py > $O/r2_bench_n1_final_full.json 2> $O/bench_x_full.err; step bench_full $?
timeout 300 python bench.py --workload pandareach --steps 200 --warmup 10 > $O/r2_bench_n1_final_pandareach.json 2> $O/bench_x_reach.err; step bench_reach $?
timeout 300 python bench.py --workload pandagrasp --steps 200 --warmup 10 > $O/r2_bench_n1_final_pandagrasp.json 2> $O/bench_x_grasp.err; step bench_grasp $?
timeout 300 compute-sanitizer --tool memcheck --print-limit 4 python tools/sanitize_case.py > $O/sanitize_x.log 2>&1; step memcheck $?
echo done >> $O/steps_x.log
tail -4 $O/pytest_x.log | cut -c1-200; cat $O/smoke_x.log; cat $O/steps_x.log; tail -3 $O/sanitize_x.log
for f in k20 full pandareach pandagrasp; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2_bench_n1_final_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.3f M"%(d["value"]/1e6), "e2e %.3f M"%(d["e2e"]["value"]/1e6), "cpu", (d.get("cpu_baseline") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"), d.get("config",{}).get("kernel_ms_by_replica"))
except Exception as e:
    print("$f failed", e)
PY
done
