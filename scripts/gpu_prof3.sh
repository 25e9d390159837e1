#!/bin/bash
# third-session profile: launch list of the bench step + ncu --set full of the f16 build group, the lookup, the fused
# encoder head and the fp16 on-the-fly lookup (each after its own command has run cleanly without ncu)
set -u
OUT=gpurun_out; TAG=${1:-q1}; mkdir -p $OUT
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fused > $OUT/${TAG}_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fused > $OUT/${TAG}_ncu1.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:build_tc_kernel|lookup_fwd|f16_operand" -s 42 -c 3 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fused > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full (bench) exit $?"
timeout 300 python scripts/fnet_head_time.py > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:fnet_head" -s 3 -c 1 -f -o $OUT/${TAG}_prof_head \
    python scripts/fnet_head_time.py > $OUT/${TAG}_ncu3.log 2>&1
echo "ncu head exit $?"
timeout 300 python scripts/alt_tc_stats.py a > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:alt_fwd_tc" -s 2 -c 1 -f -o $OUT/${TAG}_prof_alt \
    python scripts/alt_tc_stats.py a > $OUT/${TAG}_ncu4.log 2>&1
echo "ncu alt exit $?"
ls -la $OUT | grep $TAG
