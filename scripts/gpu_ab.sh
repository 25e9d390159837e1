#!/bin/bash
# gpurun payload: A/B of env-switched kernel variants on one box.  usage: gpu_ab.sh TAG "ENV=a" "ENV=b" ...
set -u
OUT=gpurun_out; TAG=${1:-ab}; shift; mkdir -p $OUT
for rep in 1 2; do
  for v in "$@"; do
    env $v timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2> $OUT/${TAG}_err.log | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'rep$rep', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
  done
done
for v in "$@"; do
  echo "== $v lookup_scaling"; env $v timeout 300 python scripts/lookup_scaling.py 2>&1 | grep "L=4 r=4\|L=4 r=3\|L=2 r=4"
  env $v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
done
