#!/bin/bash
# Round-2 call R: fewer tail blocks (denser tail: the main launch keeps more SMs).
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/steps_r.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_r_$tag.json 2> $O/bench_r_$tag.err; echo "bench $tag exit $?" >> $O/steps_r.log; }
run b64 B2ENV_TAIL_BLOCKS=64
run b32 B2ENV_TAIL_BLOCKS=32
run b24 B2ENV_TAIL_BLOCKS=24
run b16 B2ENV_TAIL_BLOCKS=16
run b32_c300 B2ENV_TAIL_BLOCKS=32 B2ENV_TAIL_COST=300000
cat $O/steps_r.log
for f in b64 b32 b24 b16 b32_c300; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_r_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_r_$f.err").read()[-800:])
PY
done
