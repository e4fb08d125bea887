#!/bin/bash
# Round-2 call P: remaining parity test detail, iCub bench with cost-ordered blocks, ncu launch list (300 steps of one rollout),
# ncu --set full of timed launches under --replicas 8 (summarised on the box: the reports exceed what gpurun_out/ may carry).
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_p.log; }
rm -f $O/steps_p.log
timeout 600 python -m pytest tests -m gpu -q --tb=short > $O/pytest_p.log 2>&1; step pytest $?
timeout 300 python bench.py --workload icubpush --steps 200 --warmup 10 --replicas 4 --no-cpu-baseline > $O/bench_p_icub.json 2> $O/bench_p_icub.err; step bench_icub $?
B2ENV_SCHED=0 timeout 300 python bench.py --workload icubpush --steps 200 --warmup 10 --replicas 4 --no-cpu-baseline > $O/bench_p_icub_nosched.json 2> $O/bench_p_icub_nosched.err; step bench_icub_nosched $?
PB1="python bench.py --replicas 1 --steps 300 --warmup 50 --no-cpu-baseline --e2e-steps 8"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_kernel -c 720 --csv --log-file $O/r2_panda_launches.csv $PB1 > $O/ncu_p_list.log 2>&1; step ncu_list $?
PB8="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 8"
# launches before the timed region: 3 settle steps + pre-roll 4383 steps + 5 warm-up steps, TWO launches (tail + main) per step
timeout 400 ncu --set full --clock-control none -k regex:step_kernel -s 8782 -c 16 -f -o /tmp/r2_panda_timed $PB8 > $O/ncu_p_timed.log 2>&1; step ncu_timed $?
python tools/ncu_multi_summary.py /tmp/r2_panda_timed.ncu-rep $O/r2_ncu_step_kernel_timed16.csv "bench.py --steps 20 --warmup 5 (8 replicas at protocol depths 50..1050): 16 consecutive launches of the timed region = 8 steps x (tail launch, main launch); ncu --set full --clock-control none" > /dev/null 2>&1; step summary $?
ls -la /tmp/r2_panda_timed.ncu-rep >> $O/steps_p.log
echo done >> $O/steps_p.log
tail -40 $O/pytest_p.log | cut -c1-200; cat $O/steps_p.log
for f in icub icub_nosched; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_p_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.3f M"%(d["value"]/1e6), "e2e %.3f M"%(d["e2e"]["value"]/1e6), d.get("config",{}).get("kernel_ms_by_replica"))
except Exception as e:
    print("$f failed", e); print(open("$O/bench_p_$f.err").read()[-1200:])
PY
done
head -5 $O/r2_ncu_step_kernel_timed16.csv | cut -c1-300
