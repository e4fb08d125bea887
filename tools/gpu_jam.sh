#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-jam}
B2ENV_SCHED=0 timeout 200 python tools/jam_profile.py 900 148 > $O/${TAG}_plain.log 2>&1; echo "plain exit $?" > $O/steps_${TAG}.log
B2ENV_SCHED=0 timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o $O/${TAG}_full python tools/jam_profile.py 900 148 > $O/${TAG}_ncu.log 2>&1; echo "ncu exit $?" >> $O/steps_${TAG}.log
cat $O/steps_${TAG}.log; cat $O/${TAG}_plain.log | tail -8
