#!/bin/bash
# round-2 single-GPU check: the whole -m gpu suite, smoke(), the default bench line, batch-1 step timings
set -u
OUT=gpurun_out; TAG=${1:-f1}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cut -c1-300 $OUT/${TAG}_bench.json
timeout 200 python scripts/lookup_time.py 1 128 46 62 3 > $OUT/${TAG}_b1_cfg1.txt 2>&1; tail -2 $OUT/${TAG}_b1_cfg1.txt
timeout 200 python scripts/lookup_time.py 1 > $OUT/${TAG}_b1_cfg2.txt 2>&1; tail -2 $OUT/${TAG}_b1_cfg2.txt
