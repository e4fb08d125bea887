#!/bin/bash
# Short gpurun call for the iCub tree kernel: its GPU tests, the iCub bench line, a launch list and one full capture.
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-icub}
timeout 300 python -m pytest tests/test_gpu_icub.py -q > $O/pytest_${TAG}.log 2>&1; echo "pytest exit $?" > $O/steps_${TAG}.log
timeout 240 python bench.py --workload icubpush --batch 16384 --steps 800 --warmup 80 --replicas 4 --cpu-batch 1024 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err; echo "bench exit $?" >> $O/steps_${TAG}.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:tree_step -s 16 -c 1 -f -o $O/${TAG}_full python bench.py --workload icubpush --batch 16384 --steps 24 --warmup 8 --replicas 4 --no-cpu-baseline --e2e-steps 8 > $O/ncu_${TAG}_full.log 2>&1; echo "ncu full exit $?" >> $O/steps_${TAG}.log
tail -3 $O/pytest_${TAG}.log; cat $O/steps_${TAG}.log; cut -c1-400 $O/bench_${TAG}.json
