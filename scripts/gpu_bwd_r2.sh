#!/bin/bash
# round-2: batched lookup adjoint -- parity tests, then the config-3 training step both ways (CUDA graph + eager) and an
# ncu launch list of one graph-free step
set -u
OUT=gpurun_out; TAG=${1:-bw1}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_lookup_bwd_batched.py tests/test_gpu_grad_train_shape.py -q -m gpu -x -p no:cacheprovider -s > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|error|dfmap" $OUT/${TAG}_pytest.log | tail -30
for MODE in batched immediate; do
  RAFTCORR_LOOKUP_BWD=$MODE timeout 300 python scripts/train_step_graph.py > $OUT/${TAG}_graph_$MODE.txt 2>&1; tail -1 $OUT/${TAG}_graph_$MODE.txt
  RAFTCORR_LOOKUP_BWD=$MODE timeout 300 python scripts/bench_configs.py 3 > $OUT/${TAG}_cfg3_$MODE.jsonl 2>&1; head -1 $OUT/${TAG}_cfg3_$MODE.jsonl | cut -c1-300
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"lookup_bwd|build_bwd|pool_adjoint|round_pad|build_tc|lookup_fwd" -c 60 --csv --log-file $OUT/${TAG}_launches.csv python scripts/train_step_probe.py > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
