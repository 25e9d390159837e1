#!/bin/bash
# round-2 multi-GPU session (one 8-GPU box): NCCL tests, headline bench and the other BASELINE.json configs at 1/2/4/8
# GPUs, the config-3 training step with the NCCL gradient all-reduce.  usage: gpu_multi_r2.sh [tag] [Ns...]
set -u
OUT=gpurun_out; TAG=${1:-r02}; shift || true; NS=${@:-1 2 4 8}; mkdir -p $OUT
nvidia-smi -L | wc -l > $OUT/${TAG}_ngpus.txt
timeout 600 python -m pytest tests/test_gpu_ddp_nccl.py tests/test_gpu_threads.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_mg_pytest.log 2>&1
tail -2 $OUT/${TAG}_mg_pytest.log
for N in $NS; do
  if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N"; fi
  timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 3 --no-fused --no-gpu-baseline --no-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
  echo "bench N=$N exit $?"; cut -c1-200 $OUT/${TAG}_bench_n$N.json
  timeout 900 $RUN scripts/bench_configs.py 1 3 4 5a 5b > $OUT/${TAG}_configs_n$N.jsonl 2> $OUT/${TAG}_configs_n$N.err
  echo "configs N=$N exit $?"
  timeout 600 $RUN scripts/train_step_ddp.py > $OUT/${TAG}_train_ddp_n$N.json 2> $OUT/${TAG}_train_ddp_n$N.err
  echo "train N=$N exit $?"; cut -c1-300 $OUT/${TAG}_train_ddp_n$N.json
done
