#!/bin/bash
# build kernel A/B: usage scripts/gpu_build_ab.sh <tag> "ENV1=.. ENV2=.." "ENV.." ...   (first arg set may be empty "")
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for envs in "$@"; do
  echo "== [$envs]" | tee -a $OUT/${TAG}_ab.log
  env $envs python scripts/tc_debug.py 2>&1 | tail -3 | cut -c1-200 | tee -a $OUT/${TAG}_ab.log
  env $envs python scripts/build_time.py 256 128 2>&1 | tee -a $OUT/${TAG}_ab.log
done
