#!/bin/bash
# Round-2 final single-GPU call: GPU tests, smoke, memcheck, bench lines of every workload (with CPU arms), reference arm,
# ncu launch list along a full rollout, ncu --set full of timed launches under --replicas 8 (tail + main launches at 8 depths).
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_final.log; }
rm -f $O/steps_final.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi_final.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_final.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; step smoke $?
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_final_k20.json 2> $O/bench_final_k20.err; step bench_k20 $?
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_final_ref.json 2> $O/bench_final_ref.err; step bench_ref $?
timeout 400 python bench.py > $O/bench_final_full.json 2> $O/bench_final_full.err; step bench_full $?
timeout 300 python bench.py --workload pandareach --steps 200 --warmup 10 > $O/bench_final_reach.json 2> $O/bench_final_reach.err; step bench_reach $?
timeout 300 python bench.py --workload pandagrasp --steps 200 --warmup 10 > $O/bench_final_grasp.json 2> $O/bench_final_grasp.err; step bench_grasp $?
timeout 300 python bench.py --workload icubpush --steps 200 --warmup 10 --replicas 4 > $O/bench_final_icub.json 2> $O/bench_final_icub.err; step bench_icub $?
PB1="python bench.py --replicas 1 --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 8"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_kernel -c 2200 --csv --log-file $O/final_panda_launches.csv $PB1 > $O/ncu_final_list.log 2>&1; step ncu_list $?
PB8="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 8"
# launches before the timed region: 3 settle steps + pre-roll 4383 steps + 5 warm-up steps, TWO launches (tail + main) per step
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 8776 -c 24 -f -o $O/final_panda_timed $PB8 > $O/ncu_final_timed.log 2>&1; step ncu_timed $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tree_step -s 120 -c 1 -f -o $O/final_icub_full python bench.py --workload icubpush --batch 16384 --steps 200 --warmup 8 --replicas 1 --no-cpu-baseline --e2e-steps 8 > $O/ncu_final_icub.log 2>&1; step ncu_icub $?
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_case.py > $O/sanitize_final.log 2>&1; step memcheck $?
echo done >> $O/steps_final.log
tail -6 $O/pytest_final.log; cat $O/smoke_final.log; cat $O/steps_final.log; tail -12 $O/sanitize_final.log
for f in k20 ref full reach grasp icub; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_final_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.3f M"%(d["value"]/1e6), "e2e %.3f M"%(d["e2e"]["value"]/1e6), "cpu", (d.get("cpu_baseline") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"), d.get("config",{}).get("kernel_ms_by_replica"))
except Exception as e:
    print("$f failed", e); print(open("$O/bench_final_$f.err").read()[-1200:])
PY
done
