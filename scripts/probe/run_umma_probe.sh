#!/bin/bash
# builds the probe from its source (binaries are not tracked)
P=scripts/probe/umma_mn_probe
[ -x $P ] || nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o $P scripts/probe/umma_mn_probe.cu -lcuda || exit 1
for v in 16 32 33 34 35 36 37 64 66 1; do timeout 20 $P $v 2>&1 | tail -1; done
