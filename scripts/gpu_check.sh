#!/bin/bash
# One gpurun call: parity tests, bench, ncu launch list + one full capture of the lookup kernel.
# usage: scripts/gpu_check.sh <tag> <mode: fp32|tf32> [pytest -k expr]
set -u
TAG=${1:-r1}; MODE=${2:-fp32}; KEXPR=${3:-}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -k "$KEXPR" > $OUT/${TAG}_pytest.log 2>&1
else
  timeout 1200 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
fi
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --mode $MODE --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
timeout 300 python bench.py --mode $MODE --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --mode $MODE --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
timeout 300 python bench.py --mode $MODE --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lookup_fwd -s 12 -c 2 -f -o $OUT/${TAG}_lookup \
    python bench.py --mode $MODE --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
