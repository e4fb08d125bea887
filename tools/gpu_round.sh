#!/bin/bash
# Round-end gpurun call: GPU tests, smoke, headline bench + iCub bench (with CPU arms), ncu launch list along a full
# rollout, full ncu captures (Panda shallow + deep, iCub), sanitizer.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps.log; }
rm -f $O/steps.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; step smoke $?
timeout 200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; step bench $?
timeout 150 python bench.py --workload icubpush --batch 16384 --steps 800 --warmup 80 --replicas 4 --cpu-batch 1024 > $O/bench_icub.json 2> $O/bench_icub.err; step bench_icub $?
PB="python bench.py --replicas 1 --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 8"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_kernel -c 1060 --csv --log-file $O/panda_launches.csv $PB > $O/ncu_panda_list.log 2>&1; step ncu_list $?
timeout 90 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -f -o $O/panda_shallow_full $PB > $O/ncu_panda_shallow.log 2>&1; step ncu_shallow $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1000 -c 1 -f -o $O/panda_deep_full $PB > $O/ncu_panda_deep.log 2>&1; step ncu_deep $?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:tree_step -s 120 -c 1 -f -o $O/icub_full python bench.py --workload icubpush --batch 16384 --steps 200 --warmup 8 --replicas 1 --no-cpu-baseline --e2e-steps 8 > $O/ncu_icub_full.log 2>&1; step ncu_icub $?
timeout 100 compute-sanitizer --tool memcheck python tools/sanitize_case.py > $O/sanitize_mem.log 2>&1; step memcheck $?
echo done >> $O/steps.log
tail -4 $O/pytest_gpu.log; cat $O/smoke.log; cat $O/steps.log; cut -c1-300 $O/bench_n1.json; echo; cut -c1-300 $O/bench_icub.json
