#!/bin/bash
# Round-2 call Y: Delassus block built from 16-byte loads: bench + stage stamps of the slowest solves.
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_y.json 2> $O/bench_y.err; echo "bench exit $?"
B2ENV_LIB=$PWD/variants/libb2env_sweep.so timeout 300 python tools/stage_profile.py 50,1000 > $O/stages_y.log 2>&1
grep -E "stage cycles|launch|solve: |collision  " $O/stages_y.log | cut -c1-300
python - <<PY
import json
d=json.loads(open("$O/bench_y.json").read().strip().splitlines()[-1])
print("value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"])
PY
