#!/bin/bash
# Round-2 call F: tail launch shapes (blocks x warps, spread slot mapping, SM-exclusive shared-memory request).
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_f.log; }
rm -f $O/steps_f.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_deep.py -m gpu -q -x > $O/pytest_f.log 2>&1; step pytest $?
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 80 --warmup 8 --no-cpu-baseline > $O/bench_f_$tag.json 2> $O/bench_f_$tag.err; step bench_$tag $?; }
run w1_b512_x0 B2ENV_TAIL_WPB=1 B2ENV_TAIL_BLOCKS=512 B2ENV_TAIL_EXCL=0
run w1_b64_x1 B2ENV_TAIL_WPB=1 B2ENV_TAIL_BLOCKS=64 B2ENV_TAIL_EXCL=1
run w4_b8 B2ENV_TAIL_WPB=4 B2ENV_TAIL_BLOCKS=8
run w4_b16 B2ENV_TAIL_WPB=4 B2ENV_TAIL_BLOCKS=16
run w4_b32 B2ENV_TAIL_WPB=4 B2ENV_TAIL_BLOCKS=32
run w2_b32_x1 B2ENV_TAIL_WPB=2 B2ENV_TAIL_BLOCKS=32 B2ENV_TAIL_EXCL=1
run w4_b16_c80 B2ENV_TAIL_WPB=4 B2ENV_TAIL_BLOCKS=16 B2ENV_TAIL_COST=80000
run w4_b16_c300 B2ENV_TAIL_WPB=4 B2ENV_TAIL_BLOCKS=16 B2ENV_TAIL_COST=300000
echo done >> $O/steps_f.log
tail -5 $O/pytest_f.log; cat $O/steps_f.log
for f in w1_b512_x0 w1_b64_x1 w4_b8 w4_b16 w4_b32 w2_b32_x1 w4_b16_c80 w4_b16_c300; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_f_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], d["config"]["mean_pgs_iters_last_step"], d["config"]["sweep_capped_envs_last_step"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_f_$f.err").read()[-1500:])
PY
done
