#!/bin/bash
# Round-2 call U: the motor block of coupled systems — fully unrolled block form (MOTOR_UNROLL=9), serial rows only.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/steps_u.log
run() { tag=$1; lib=$2; shift; shift; env B2ENV_LIB=$lib "$@" timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_u_$tag.json 2> $O/bench_u_$tag.err; echo "bench $tag exit $?" >> $O/steps_u.log; }
run base ""
run u9 $PWD/variants/libb2env_u9.so
run u9_w1 $PWD/variants/libb2env_u9.so B2ENV_TAIL_WPB=1 B2ENV_TAIL_BLOCKS=256
run ser $PWD/variants/libb2env_ser.so
B2ENV_LIB=$PWD/variants/libb2env_sweep.so timeout 300 python tools/stage_profile.py 1000 > $O/stages_u.log 2>&1
cat $O/steps_u.log; grep -E "sweep parts|before the serial|launch" $O/stages_u.log | cut -c1-230
for f in base u9 u9_w1 ser; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_u_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"])
except Exception as e:
    print("$f failed", e); print(open("$O/bench_u_$f.err").read()[-800:])
PY
done
