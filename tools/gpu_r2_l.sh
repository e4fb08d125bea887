#!/bin/bash
# Round-2 call L: compact rolled sweep (one row loop in the code for systems of more than 16 generic rows), row loop unrolled 1 / 2 / 4;
# parity tests with the restated sweep-capped criteria.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/steps_l.log
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_deep.py -m gpu -q > $O/pytest_l.log 2>&1; echo "pytest exit $?" >> $O/steps_l.log
for v in base u2 u4; do
  if [ "$v" = "base" ]; then L=""; else L="$PWD/variants/libb2env_$v.so"; fi
  B2ENV_LIB=$L B2ENV_SCHED=0 timeout 200 python tools/jam_profile.py 900 148 > $O/jam_l_$v.log 2>&1; echo "jam $v exit $?" >> $O/steps_l.log
  B2ENV_LIB=$L timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_l_$v.json 2> $O/bench_l_$v.err; echo "bench $v exit $?" >> $O/steps_l.log
done
tail -6 $O/pytest_l.log; cat $O/steps_l.log
for v in base u2 u4; do tail -2 $O/jam_l_$v.log; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_l_$v.json").read().strip().splitlines()[-1])
    print("$v", "value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"])
except Exception as e:
    print("$v failed", e); print(open("$O/bench_l_$v.err").read()[-800:])
PY
done
