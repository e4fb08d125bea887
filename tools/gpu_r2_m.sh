#!/bin/bash
# Round-2 call M: fine stage profile (parts of `solve` and `rest`).
mkdir -p gpurun_out
O=gpurun_out
B2ENV_LIB=$PWD/variants/libb2env_stages.so timeout 300 python tools/stage_profile.py 50,600 > $O/stages_m.log 2>&1; echo "stages exit $?"
cut -c1-170 $O/stages_m.log
