#!/bin/bash
# Round-2 call H: failing parity tests again (tie tolerances), jam population GPU vs CPU, bench with tail shapes for the heavy-env regime.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_h.log; }
rm -f $O/steps_h.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_deep.py -m gpu -q > $O/pytest_h.log 2>&1; step pytest $?
timeout 600 python tools/jam_compare.py 4096 1000 > $O/jam_compare.log 2>&1; step jam_compare $?
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_h_$tag.json 2> $O/bench_h_$tag.err; step bench_$tag $?; }
run default
run w4_b128_x0 B2ENV_TAIL_WPB=4 B2ENV_TAIL_BLOCKS=128 B2ENV_TAIL_EXCL=0
run w2_b256_x0 B2ENV_TAIL_WPB=2 B2ENV_TAIL_BLOCKS=256 B2ENV_TAIL_EXCL=0
run w1_b512_x0 B2ENV_TAIL_WPB=1 B2ENV_TAIL_BLOCKS=512 B2ENV_TAIL_EXCL=0
run nosched B2ENV_SCHED=0
echo done >> $O/steps_h.log
tail -12 $O/pytest_h.log; cat $O/steps_h.log; cat $O/jam_compare.log | cut -c1-400
for f in default w4_b128_x0 w2_b256_x0 w1_b512_x0 nosched; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_h_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], d["config"]["mean_pgs_iters_last_step"], d["config"]["sweep_capped_envs_last_step"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_h_$f.err").read()[-1500:])
PY
done
