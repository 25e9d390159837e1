#!/bin/bash
# gpurun payload: bench-only A/B of env-switched variants.  usage: gpu_ab2.sh TAG "ENV=a" "ENV=b" ...
set -u
OUT=gpurun_out; TAG=${1:-ab}; shift; mkdir -p $OUT
for rep in 1 2; do
  for v in "$@"; do
    env $v timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2> $OUT/${TAG}_err.log | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$v', 'rep$rep', round(d['value'],1), round(d['ms_per_step'],4), 'build', round(k['build']['ms_per_launch_group'],4), 'lookup', round(k['lookup']['ms_per_launch'],5))"
  done
done
