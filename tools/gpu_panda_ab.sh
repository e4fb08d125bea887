#!/bin/bash
# One gpurun call: all GPU tests, then the headline bench with slow-first scheduling on (default) and off.
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-ab}
timeout 500 python -m pytest tests -m gpu -q > $O/pytest_${TAG}.log 2>&1; echo "pytest exit $?" > $O/steps_${TAG}.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_${TAG}_sched.json 2> $O/bench_${TAG}_sched.err; echo "sched exit $?" >> $O/steps_${TAG}.log
B2ENV_SCHED=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_${TAG}_nosched.json 2> $O/bench_${TAG}_nosched.err; echo "nosched exit $?" >> $O/steps_${TAG}.log
tail -3 $O/pytest_${TAG}.log; cat $O/steps_${TAG}.log
for f in $O/bench_${TAG}_*.json; do echo "$f $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('kernel_ms_at_final_depth'))" 2>&1 | tail -1)"; done
