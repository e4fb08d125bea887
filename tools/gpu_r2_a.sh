#!/bin/bash
# Round-2 call A: GPU tests + the depth-invariant bench at the driver's K (20) and at the full protocol, the CPU arm,
# and ncu captures of 8 consecutive timed launches (one per replica = one per rollout depth) under --replicas 8.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "$1 exit $2 t=$(( $(date +%s)-T0 ))" >> $O/steps_a.log; }
rm -f $O/steps_a.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 400 python -m pytest tests -m gpu -q -x > $O/pytest_a.log 2>&1; step pytest $?
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_a.log 2>&1; step smoke $?
timeout 200 python bench.py --steps 20 --warmup 5 > $O/bench_a_k20.json 2> $O/bench_a_k20.err; step bench_k20 $?
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_a_k20b.json 2> $O/bench_a_k20b.err; step bench_k20b $?
timeout 300 python bench.py > $O/bench_a_full.json 2> $O/bench_a_full.err; step bench_full $?
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_a_ref.json 2> $O/bench_a_ref.err; step bench_ref $?
PB="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 8"
# launches before the timed region: 3 settle launches + pre-roll (plan: 49+191+334+476+619+762+905+1047 = 4383) + 5 warm-up
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4391 -c 8 -f -o $O/r2a_panda_timed8 $PB > $O/ncu_a_timed8.log 2>&1; step ncu_timed8 $?
echo done >> $O/steps_a.log
tail -4 $O/pytest_a.log; cat $O/smoke_a.log; cat $O/steps_a.log; for f in k20 k20b full ref; do cut -c1-400 $O/bench_a_$f.json; echo; done
