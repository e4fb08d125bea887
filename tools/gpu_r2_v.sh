#!/bin/bash
# Round-2 call V: serial motor rows only + staged friction bounds: bench, sweep-part profile, parity tests.
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_v.json 2> $O/bench_v.err; echo "bench exit $?"
B2ENV_LIB=$PWD/variants/libb2env_sweep.so timeout 300 python tools/stage_profile.py 1000 > $O/stages_v.log 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_deep.py -m gpu -q --tb=short > $O/pytest_v.log 2>&1; echo "pytest exit $?"
tail -3 $O/pytest_v.log
grep -E "sweep parts|before the serial|launch|solve: " $O/stages_v.log | cut -c1-230
python - <<PY
import json
d=json.loads(open("$O/bench_v.json").read().strip().splitlines()[-1])
print("value %.2f M"%(d["value"]/1e6), "e2e %.2f M"%(d["e2e"]["value"]/1e6), d["config"]["kernel_ms_by_replica"])
PY
