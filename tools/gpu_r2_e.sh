#!/bin/bash
# Round-2 call E: look-ahead tail solver A/B (B2ENV_TAIL_LA, B2ENV_TAIL_COST, B2ENV_TAIL_WPB), microbenchmark row_chain3.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_e.log; }
rm -f $O/steps_e.log
(cd tools/micro && timeout 60 ./row_chain3 > ../../$O/row_chain3.log 2>&1); step micro $?
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_deep.py -m gpu -q -x > $O/pytest_e.log 2>&1; step pytest $?
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 80 --warmup 8 --no-cpu-baseline > $O/bench_e_$tag.json 2> $O/bench_e_$tag.err; step bench_$tag $?; }
run la0 B2ENV_TAIL_LA=0
run la1 B2ENV_TAIL_LA=1
run la1_c80 B2ENV_TAIL_LA=1 B2ENV_TAIL_COST=80000
run la1_c40 B2ENV_TAIL_LA=1 B2ENV_TAIL_COST=40000
run la1_c40_w2 B2ENV_TAIL_LA=1 B2ENV_TAIL_COST=40000 B2ENV_TAIL_WPB=2
run la1_c20 B2ENV_TAIL_LA=1 B2ENV_TAIL_COST=20000
echo done >> $O/steps_e.log
cat $O/row_chain3.log; tail -5 $O/pytest_e.log; cat $O/steps_e.log
for f in la0 la1 la1_c80 la1_c40 la1_c40_w2 la1_c20; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_e_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], d["config"]["mean_pgs_iters_last_step"], d["config"]["sweep_capped_envs_last_step"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_e_$f.err").read()[-1500:])
PY
done
