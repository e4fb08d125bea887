#!/bin/bash
# Round-2 call: GPU tests + bench at K=20 for scheduling variants (B2ENV_SCHED=0, tail block sizes) + stage profile.
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-b}
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_$TAG.log; }
rm -f $O/steps_$TAG.log
timeout 400 python -m pytest tests -m gpu -q -x > $O/pytest_$TAG.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; step smoke $?
for v in 4 2 1; do
  B2ENV_TAIL_WPB=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_tw$v.json 2> $O/bench_${TAG}_tw$v.err; step bench_tw$v $?
done
B2ENV_SCHED=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_nosched.json 2> $O/bench_${TAG}_nosched.err; step bench_nosched $?
B2ENV_LIB=$PWD/variants/libb2env_stages.so timeout 300 python tools/stage_profile.py 50,300,600,1000 > $O/stages_$TAG.log 2>&1; step stages $?
echo done >> $O/steps_$TAG.log
tail -4 $O/pytest_$TAG.log; cat $O/smoke_$TAG.log; cat $O/steps_$TAG.log
for f in tw4 tw2 tw1 nosched; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], "capped", d["config"]["sweep_capped_envs_last_step"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_${TAG}_$f.err").read()[-800:])
PY
done
cat $O/stages_$TAG.log
