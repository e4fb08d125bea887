#!/bin/bash
# Round-2 call G: all GPU tests with the completed N1 model (pads, capsule, self pairs), smoke, bench K=20 / full, stage profile.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_g.log; }
rm -f $O/steps_g.log
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_g.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_g.log 2>&1; step smoke $?
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_g_k20.json 2> $O/bench_g_k20.err; step bench_k20 $?
timeout 300 python bench.py --no-cpu-baseline > $O/bench_g_full.json 2> $O/bench_g_full.err; step bench_full $?
B2ENV_LIB=$PWD/variants/libb2env_stages.so timeout 300 python tools/stage_profile.py 50,1000 > $O/stages_g2.log 2>&1; step stages $?
echo done >> $O/steps_g.log
tail -30 $O/pytest_g.log; cat $O/smoke_g.log; cat $O/steps_g.log; cat $O/stages_g2.log | cut -c1-200
for f in k20 full; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_g_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], d["config"]["mean_pgs_iters_last_step"], d["config"]["sweep_capped_envs_last_step"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_g_$f.err").read()[-1500:])
PY
done
