#!/bin/bash
# One gpurun call: iCub GPU tests with the in-tree library, then the iCub bench line for the in-tree library and for
# every variant library under variants/ (built with different -D switches; B2ENV_LIB selects the library).
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-var}
timeout 300 python -m pytest tests/test_gpu_icub.py -q > $O/pytest_${TAG}.log 2>&1; echo "pytest exit $?" > $O/steps_${TAG}.log
BENCH="python bench.py --workload icubpush --batch 16384 --steps 600 --warmup 80 --replicas 4 --no-cpu-baseline --e2e-steps 8"
timeout 120 $BENCH > $O/bench_${TAG}_main.json 2> $O/bench_${TAG}_main.err; echo "main exit $?" >> $O/steps_${TAG}.log
for v in variants/*.so; do
  n=$(basename $v .so)
  B2ENV_LIB=$PWD/$v timeout 120 $BENCH > $O/bench_${TAG}_$n.json 2> $O/bench_${TAG}_$n.err; echo "$n exit $?" >> $O/steps_${TAG}.log
done
tail -2 $O/pytest_${TAG}.log; cat $O/steps_${TAG}.log
for f in $O/bench_${TAG}_*.json; do echo "$f $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'])" 2>&1 | tail -1)"; done
