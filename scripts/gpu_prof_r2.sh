#!/bin/bash
# round-2 profile: ncu --set full of the lookup, the f16 build group and the on-the-fly lookup (after a clean plain run)
set -u
OUT=gpurun_out; TAG=${1:-r2a}; KER=${2:-"regex:lookup_fwd_tma|build_tc_kernel|alt_fwd_tc|f16_operand"}; ARGS=${3:-alt}; mkdir -p $OUT
timeout 300 python scripts/prof_driver.py $ARGS > $OUT/${TAG}_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k "$KER" -c ${4:-12} -f -o $OUT/${TAG}_prof \
    python scripts/prof_driver.py $ARGS > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/${TAG}_ncu.log
