#!/bin/bash
# Round-2 call I: parity tests (restated sweep-capped criteria), bench after the collision-stage votes, jammed-solve profile (ncu full).
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_i.log; }
rm -f $O/steps_i.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_deep.py tests/test_gpu_ik_goal.py -m gpu -q > $O/pytest_i.log 2>&1; step pytest $?
timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_i_k40.json 2> $O/bench_i_k40.err; step bench $?
B2ENV_SCHED=0 timeout 200 python tools/jam_profile.py 900 148 > $O/jam_i_plain.log 2>&1; step jam_plain $?
B2ENV_SCHED=0 timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o $O/jam_i_full python tools/jam_profile.py 900 148 > $O/jam_i_ncu.log 2>&1; step jam_ncu $?
echo done >> $O/steps_i.log
tail -12 $O/pytest_i.log; cat $O/steps_i.log; tail -8 $O/jam_i_plain.log
python - <<PY
import json
d=json.loads(open("$O/bench_i_k40.json").read().strip().splitlines()[-1])
print("k40", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], d["config"]["mean_pgs_iters_last_step"], d["config"]["sweep_capped_envs_last_step"])
PY
