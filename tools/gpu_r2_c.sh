#!/bin/bash
# Round-2 call C: GPU tests, headline bench (K=20 and full protocol), the other BASELINE workloads on one GPU.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_c.log; }
rm -f $O/steps_c.log
timeout 400 python -m pytest tests -m gpu -q -x > $O/pytest_c.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_c.log 2>&1; step smoke $?
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c_k20.json 2> $O/bench_c_k20.err; step bench_k20 $?
timeout 300 python bench.py --no-cpu-baseline > $O/bench_c_full.json 2> $O/bench_c_full.err; step bench_full $?
B2ENV_SCHED=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_c_full_nosched.json 2> $O/bench_c_full_nosched.err; step bench_full_nosched $?
timeout 300 python bench.py --workload pandareach --steps 200 --warmup 10 > $O/bench_c_reach.json 2> $O/bench_c_reach.err; step bench_reach $?
timeout 300 python bench.py --workload pandagrasp --steps 200 --warmup 10 > $O/bench_c_grasp.json 2> $O/bench_c_grasp.err; step bench_grasp $?
timeout 300 python bench.py --workload icubpush --steps 200 --warmup 10 --replicas 4 > $O/bench_c_icub.json 2> $O/bench_c_icub.err; step bench_icub $?
echo done >> $O/steps_c.log
tail -3 $O/pytest_c.log; cat $O/smoke_c.log; cat $O/steps_c.log
for f in k20 full full_nosched reach grasp icub; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_c_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), "cpu", d.get("cpu_baseline",{}).get("value"), d["config"]["kernel_ms_by_replica"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_c_$f.err").read()[-1500:])
PY
done
