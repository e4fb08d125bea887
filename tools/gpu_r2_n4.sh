#!/bin/bash
# Round-2 4-GPU call: strong scaling of the headline (16384 envs in total) at N = 4 and N = 2 (two of the four GPUs).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/steps_n4.log
for N in 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 8 --no-cpu-baseline --e2e-steps 16 --workload pandapush --scaling strong --batch 16384 > $O/r2_bench_n${N}_pandapush_strong_16384.json 2> $O/bench_n${N}_push_strong.err; echo "strong $N exit $?" >> $O/steps_n4.log
done
cat $O/steps_n4.log
for N in 4 2; do python - <<PY
import json
try:
    d=json.loads(open("$O/r2_bench_n${N}_pandapush_strong_16384.json").read().strip().splitlines()[-1])
    print("N=$N", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["n_gpus"], d["scaling"], d["config"]["global_batch"], d["config"]["kernel_ms_by_replica"])
except Exception as e:
    print("N=$N failed", e); print(open("$O/bench_n${N}_push_strong.err").read()[-1200:])
PY
done
