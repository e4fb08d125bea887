#!/bin/bash
# Round-2 call W: pointer-increment motor loop; row loops unrolled by 2 again now that the block form is gone.
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_w.json 2> $O/bench_w.err; echo "bench exit $?"
B2ENV_LIB=$PWD/variants/libb2env_u2.so timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_w_u2.json 2> $O/bench_w_u2.err; echo "bench u2 exit $?"
B2ENV_LIB=$PWD/variants/libb2env_sweep.so timeout 300 python tools/stage_profile.py 1000 > $O/stages_w.log 2>&1
grep -E "sweep parts|launch" $O/stages_w.log | cut -c1-230
for f in w w_u2; do python - <<PY
import json
d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
print("$f value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"])
PY
done
