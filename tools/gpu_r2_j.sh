#!/bin/bash
# Round-2 call J: solver loop trimming (21-instruction row loop, delta-form motor rows, sticky serial path): tests, bench, jam profile,
# stage profile (collision stage after the votes).
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_j.log; }
rm -f $O/steps_j.log
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_j.log 2>&1; step pytest $?
timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_j_k40.json 2> $O/bench_j_k40.err; step bench $?
B2ENV_SCHED=0 timeout 200 python tools/jam_profile.py 900 148 > $O/jam_j_plain.log 2>&1; step jam_plain $?
B2ENV_LIB=$PWD/variants/libb2env_stages.so timeout 300 python tools/stage_profile.py 50,1000 > $O/stages_j.log 2>&1; step stages $?
echo done >> $O/steps_j.log
tail -12 $O/pytest_j.log; cat $O/steps_j.log; tail -5 $O/jam_j_plain.log; cut -c1-160 $O/stages_j.log
python - <<PY
import json
d=json.loads(open("$O/bench_j_k40.json").read().strip().splitlines()[-1])
print("k40", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], d["config"]["mean_pgs_iters_last_step"], d["config"]["sweep_capped_envs_last_step"])
PY
