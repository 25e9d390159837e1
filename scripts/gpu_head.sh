#!/bin/bash
# fnet head: parity tests + timing + launch list.  usage: scripts/gpu_head.sh <tag> [all]
set -u
TAG=${1:-h1}; OUT=gpurun_out; mkdir -p $OUT
if [ "${2:-}" = "all" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
else
  timeout 900 python -m pytest tests/test_gpu_fnet_head.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
fi
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -40 $OUT/${TAG}_pytest.log
timeout 300 python scripts/fnet_head_time.py > $OUT/${TAG}_head_time.jsonl 2> $OUT/${TAG}_head_time.err
echo "time exit $?"; cat $OUT/${TAG}_head_time.jsonl; tail -5 $OUT/${TAG}_head_time.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fnet_head|build_tc|nchw_to_nhwc|cudnn|cutlass|gemm|conv" -c 60 --csv \
   --log-file $OUT/${TAG}_head_launches.csv python scripts/fnet_head_time.py > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; grep -v "^==" $OUT/${TAG}_head_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -20
