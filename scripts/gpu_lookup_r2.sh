#!/bin/bash
# round-2 lookup session: parity tests of everything that reads the pyramid, then per-launch time (CUDA graph of 12) with
# the packed and the two-windows-per-pass evaluation, and the gather-only / evaluate-only ablations
set -u
OUT=gpurun_out; TAG=${1:-lk1}; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guard_bands.py tests/test_gpu_fuzz.py tests/test_gpu_e2e.py -q -m gpu -x -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
for P in 1 0; do
  RAFTCORR_LOOKUP_PACK=$P timeout 300 python scripts/lookup_time.py > $OUT/${TAG}_time_pack$P.txt 2>&1; tail -1 $OUT/${TAG}_time_pack$P.txt
done
timeout 300 python scripts/lookup_time.py 1 > $OUT/${TAG}_time_b1.txt 2>&1; tail -1 $OUT/${TAG}_time_b1.txt
timeout 300 python scripts/gather_roofline.py > $OUT/${TAG}_gather.json 2>$OUT/${TAG}_gather.err; cat $OUT/${TAG}_gather.json
