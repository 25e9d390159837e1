#!/bin/bash
# builds the probe from its source (binaries are not tracked) and runs the box/alignment sweep
P=scripts/probe/tma_probe
[ -x $P ] || nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o $P scripts/probe/tma_probe.cu -lcuda || exit 1
for args in "64 32 64 4 16 8 0 0 1" "64 32 64 4 16 8 5 3 1" "64 32 64 4 12 10 0 0 1" "64 32 64 4 12 10 5 3 2" "64 32 64 4 12 10 -3 -2 2" "24 16 24 4 12 10 20 12 3" "22 19 24 4 8 8 -5 15 0" "3 2 8 4 12 10 -4 -4 1" "64 32 64 4 12 10 100000 5 1" "64 32 64 4 8 8 3 3 1" "64 32 64 4 32 8 3 3 1"; do
  timeout 30 $P $args 2>&1 | tail -4
done
