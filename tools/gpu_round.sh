#!/bin/bash
# One gpurun call: GPU tests, headline bench, iCub bench, ncu launch lists + one full capture of the tree kernel,
# sanitizer pass.  Everything lands in gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $? t=$(( $(date +%s)-T0 ))" >> $O/steps.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $? t=$(( $(date +%s)-T0 ))" >> $O/steps.log
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $? t=$(( $(date +%s)-T0 ))" >> $O/steps.log
timeout 240 python bench.py --workload icubpush --batch 16384 --steps 800 --warmup 80 --replicas 4 --cpu-batch 1024 > $O/bench_icub.json 2> $O/bench_icub.err; echo "bench icub exit $? t=$(( $(date +%s)-T0 ))" >> $O/steps.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/icub_launches.csv python bench.py --workload icubpush --batch 16384 --steps 24 --warmup 8 --replicas 4 --no-cpu-baseline --e2e-steps 8 > $O/ncu_icub_list.log 2>&1; echo "ncu list exit $? t=$(( $(date +%s)-T0 ))" >> $O/steps.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:tree_step -s 16 -c 1 -o $O/icub_tree_full python bench.py --workload icubpush --batch 16384 --steps 24 --warmup 8 --replicas 4 --no-cpu-baseline --e2e-steps 8 > $O/ncu_icub_full.log 2>&1; echo "ncu full exit $? t=$(( $(date +%s)-T0 ))" >> $O/steps.log
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_case.py > $O/sanitize_mem.log 2>&1; echo "memcheck exit $? t=$(( $(date +%s)-T0 ))" >> $O/steps.log
echo done >> $O/steps.log
tail -5 $O/pytest_gpu.log; cat $O/steps.log; cut -c1-600 $O/bench_n1.json; cut -c1-600 $O/bench_icub.json
