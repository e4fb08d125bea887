#!/bin/bash
# headline bench line for the in-tree library and for every variant library under variants/
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-pv}
BENCH="python bench.py --no-cpu-baseline --e2e-steps 100"
timeout 200 $BENCH > $O/bench_${TAG}_main.json 2> $O/bench_${TAG}_main.err
for v in variants/*.so; do
  n=$(basename $v .so)
  B2ENV_LIB=$PWD/$v timeout 200 $BENCH > $O/bench_${TAG}_$n.json 2> $O/bench_${TAG}_$n.err
done
for f in $O/bench_${TAG}_*.json; do echo "$f $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['config'].get('kernel_ms_at_final_depth'))" 2>&1 | tail -1)"; done
