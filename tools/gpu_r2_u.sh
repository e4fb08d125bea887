#!/bin/bash
# probe builds of the warm-start region (see -DPROFILE_WARM): shuffle form and staged form, same rollout
O=gpurun_out; mkdir -p $O
for v in sweep; do
B2ENV_LIB=$PWD/variants/libb2env_$v.so timeout 300 python tools/stage_profile.py 1000 > $O/stages_u_$v.log 2>&1; echo "$v stages exit $?"
grep "stage cycles" $O/stages_u_$v.log | cut -c1-260
done
