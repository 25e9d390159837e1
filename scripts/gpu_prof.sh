#!/bin/bash
# gpurun payload: full GPU test-suite + smoke + bench (+ L2-fetch experiment) + ncu full capture of build and lookup
set -u
OUT=gpurun_out; TAG=${1:-p1}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cat $OUT/${TAG}_bench.json
RAFTCORR_PYRAMID_LAYOUT=rowmajor RAFTCORR_LOOKUP=cpasync timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_cpasync.json 2> $OUT/${TAG}_bench_cpasync.err; echo "bench(cpasync) exit $?"; cat $OUT/${TAG}_bench_cpasync.json
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:build_tc_kernel|lookup_fwd" -s 39 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"

timeout 600 python scripts/bench_configs.py > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err; echo "configs exit $?"; cat $OUT/${TAG}_configs.jsonl
timeout 300 python scripts/alt_tc_stats.py a b > $OUT/${TAG}_alt_tc_stats.txt 2>&1; echo "alt stats exit $?"; cat $OUT/${TAG}_alt_tc_stats.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:alt_fwd_tc" -s 2 -c 1 -f -o $OUT/${TAG}_prof_alt \
    python scripts/alt_tc_stats.py a > $OUT/${TAG}_ncu3.log 2>&1
echo "ncu alt exit $?"
ls -la $OUT | tail -20
