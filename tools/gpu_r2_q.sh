#!/bin/bash
# Round-2 call Q: final bench lines of every workload at HEAD (with CPU arms), reference arm, GPU tests.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_q.log; }
rm -f $O/steps_q.log
timeout 600 python -m pytest tests -m gpu -q --tb=short > $O/pytest_q.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_q.log 2>&1; step smoke $?
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1_final_k20.json 2> $O/bench_q_k20.err; step bench_k20 $?
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_n1_final_reference_arm.json 2> $O/bench_q_ref.err; step bench_ref $?
timeout 400 python bench.py > $O/r2_bench_n1_final_full.json 2> $O/bench_q_full.err; step bench_full $?
timeout 300 python bench.py --workload pandareach --steps 200 --warmup 10 > $O/r2_bench_n1_final_pandareach.json 2> $O/bench_q_reach.err; step bench_reach $?
timeout 300 python bench.py --workload pandagrasp --steps 200 --warmup 10 > $O/r2_bench_n1_final_pandagrasp.json 2> $O/bench_q_grasp.err; step bench_grasp $?
timeout 400 python bench.py --workload icubpush --steps 200 --warmup 10 --replicas 4 > $O/r2_bench_n1_final_icubpush.json 2> $O/bench_q_icub.err; step bench_icub $?
echo done >> $O/steps_q.log
tail -8 $O/pytest_q.log | cut -c1-200; cat $O/smoke_q.log; cat $O/steps_q.log
for f in k20 reference_arm full pandareach pandagrasp icubpush; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2_bench_n1_final_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.3f M"%(d["value"]/1e6), "e2e %.3f M"%(d["e2e"]["value"]/1e6), "cpu", (d.get("cpu_baseline") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"), "traffic", (d.get("roofline") or {}).get("traffic"), d.get("config",{}).get("kernel_ms_by_replica"))
except Exception as e:
    print("$f failed", e)
PY
done
