#!/bin/bash
# Round-2 call T: cycles per part of the sweeps of the slowest solves (tail launch live), -DPROFILE_STAGES -DPROFILE_SWEEP build.
mkdir -p gpurun_out
O=gpurun_out
B2ENV_LIB=$PWD/variants/libb2env_sweep.so timeout 300 python tools/stage_profile.py 1000 > $O/stages_t.log 2>&1; echo "stages exit $?"
cut -c1-260 $O/stages_t.log
