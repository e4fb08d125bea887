#!/bin/bash
# Round-2 call S: like-for-like with round 1 — the round-1 contact model (B2ENV_CONTACT_MODEL=r1) on the round-2 kernels,
# cost-ordered scheduling + tail launch on / off.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/steps_s.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_s_$tag.json 2> $O/bench_s_$tag.err; echo "bench $tag exit $?" >> $O/steps_s.log; }
run r1 B2ENV_CONTACT_MODEL=r1
run r1_nosched B2ENV_CONTACT_MODEL=r1 B2ENV_SCHED=0
cat $O/steps_s.log
for f in r1 r1_nosched; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_s_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"], d["config"]["sweep_capped_envs_last_step"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_s_$f.err").read()[-800:])
PY
done
