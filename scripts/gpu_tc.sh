#!/bin/bash
# gpurun payload: debug ladder for the tcgen05 build, then the full GPU test-suite and bench if it passes
set -u
OUT=gpurun_out; TAG=${1:-tc}; mkdir -p $OUT
timeout 300 python scripts/tc_debug.py > $OUT/${TAG}_debug.log 2>&1; echo "tc_debug exit $?"
cat $OUT/${TAG}_debug.log | tail -40
if grep -q "TC_DEBUG PASS" $OUT/${TAG}_debug.log; then
  bash scripts/gpu_check.sh $TAG tf32
fi
