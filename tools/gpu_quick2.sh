#!/bin/bash
# GPU tests + headline bench + jammed-tail timing (no ncu)
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-q}
timeout 500 python -m pytest tests -m gpu -q > $O/pytest_${TAG}.log 2>&1; echo "pytest exit $?" > $O/steps_${TAG}.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_${TAG}_panda.json 2> $O/bench_${TAG}_panda.err; echo "panda exit $?" >> $O/steps_${TAG}.log
B2ENV_SCHED=0 timeout 200 python tools/jam_profile.py 900 148 > $O/${TAG}_jam.log 2>&1; echo "jam exit $?" >> $O/steps_${TAG}.log
tail -3 $O/pytest_${TAG}.log; cat $O/steps_${TAG}.log; tail -5 $O/${TAG}_jam.log
for f in $O/bench_${TAG}_*.json; do echo "$f $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('kernel_ms_at_final_depth'), d['config'].get('mean_pgs_iters_last_step'))" 2>&1 | tail -1)"; done
