#!/bin/bash
mkdir -p gpurun_out
B2ENV_SCHED=0 B2ENV_LIB=$PWD/variants/v_prof.so timeout 200 python tools/jam_profile.py 900 148 > gpurun_out/prof_cycles.log 2>&1
tail -6 gpurun_out/prof_cycles.log
