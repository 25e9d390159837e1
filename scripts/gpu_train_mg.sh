#!/bin/bash
# config-3 training step on N GPUs, eager and as a CUDA graph, with the NCCL gradient all-reduce.  usage: gpu_train_mg.sh <tag> <N> [extra: bench|configs]
set -u
OUT=gpurun_out; TAG=${1:-tr}; N=${2:-2}; EXTRA=${3:-}; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
for G in 1 0; do
  TRAIN_GRAPH=$G timeout 300 $RUN scripts/train_step_ddp.py > $OUT/${TAG}_train_n${N}_graph$G.json 2> $OUT/${TAG}_train_n${N}_graph$G.err
  echo "train N=$N graph=$G exit $?"; cut -c1-420 $OUT/${TAG}_train_n${N}_graph$G.json
done
if [ -n "$EXTRA" ]; then
  timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 3 --no-fused --no-gpu-baseline --no-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
  echo "bench N=$N exit $?"; cut -c1-200 $OUT/${TAG}_bench_n$N.json
  timeout 600 $RUN scripts/bench_configs.py 5a 5b 3 > $OUT/${TAG}_configs_n$N.jsonl 2> $OUT/${TAG}_configs_n$N.err
  echo "configs N=$N exit $?"; cut -c1-260 $OUT/${TAG}_configs_n$N.jsonl
fi
