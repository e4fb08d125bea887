#!/bin/bash
# active-lane probes (-DPROFILE_WARM build): iCub tree kernel stage boundaries + loop sites, Panda loop sites
O=gpurun_out; mkdir -p $O
B2ENV_LIB=$PWD/variants/libb2env_lanes.so timeout 100 python tools/lane_probe.py 20,200 > $O/lanes_icub_w.log 2>&1; echo "icub exit $?"
cat $O/lanes_icub_w.log | cut -c1-200 | tail -40
B2ENV_LIB=$PWD/variants/libb2env_lanes.so timeout 60 python tools/stage_profile.py 200 > $O/lanes_panda_w.log 2>&1; echo "panda exit $?"
grep "loop site" $O/lanes_panda_w.log
