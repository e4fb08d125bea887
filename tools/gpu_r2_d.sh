#!/bin/bash
# Round-2 call D: first GPU run of the N1 collision stage — GPU tests, smoke, headline bench at K=20 and full protocol.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_d.log; }
rm -f $O/steps_d.log
timeout 500 python -m pytest tests -m gpu -q > $O/pytest_d.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_d.log 2>&1; step smoke $?
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_d_k20.json 2> $O/bench_d_k20.err; step bench_k20 $?
timeout 300 python bench.py --no-cpu-baseline > $O/bench_d_full.json 2> $O/bench_d_full.err; step bench_full $?
timeout 300 python bench.py --workload pandagrasp --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_d_grasp.json 2> $O/bench_d_grasp.err; step bench_grasp $?
echo done >> $O/steps_d.log
tail -15 $O/pytest_d.log; cat $O/smoke_d.log; cat $O/steps_d.log
for f in k20 full grasp; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_d_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], d["config"]["mean_pgs_iters_last_step"], d["config"]["sweep_capped_envs_last_step"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_d_$f.err").read()[-1500:])
PY
done
