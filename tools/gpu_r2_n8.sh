#!/bin/bash
# Round-2 multi-GPU call (N = number of GPUs of the box, passed as $1): BASELINE configs 4 and 5 and the strong-scaling split of config 3.
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_n$N.log; }
rm -f $O/steps_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
COMMON="--gpus $N --steps 40 --warmup 8 --no-cpu-baseline --e2e-steps 16"
timeout 300 $TR bench.py $COMMON --workload pandapush --scaling strong --batch 16384 > $O/bench_n${N}_push_strong.json 2> $O/bench_n${N}_push_strong.err; step push_strong $?
if [ "$N" = "8" ]; then
  timeout 300 $TR bench.py $COMMON --workload icubpush --batch 8192 --replicas 4 > $O/bench_n${N}_icub65536.json 2> $O/bench_n${N}_icub65536.err; step icub $?
  timeout 300 $TR bench.py $COMMON --workload pandagrasp --batch 16384 > $O/bench_n${N}_grasp131072.json 2> $O/bench_n${N}_grasp131072.err; step grasp $?
fi
echo done >> $O/steps_n$N.log
cat $O/steps_n$N.log
for f in $O/bench_n${N}_*.json; do python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), "n_gpus", d["n_gpus"], d["scaling"], d["config"]["global_batch"], d["config"]["kernel_ms_by_replica"])
except Exception as e:
    print("$f failed", e); print(open("$f".replace(".json",".err")).read()[-1200:])
PY
done
