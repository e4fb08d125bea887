#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_case.py > $O/sanitize_mem.log 2>&1; echo "memcheck exit $?"
head -80 $O/sanitize_mem.log | cut -c1-220
