#!/bin/bash
# round-2 multi-GPU session on one N-GPU box (gpurun --gpus N): NCCL tests, headline bench, the other BASELINE.json
# configs under torchrun (weak scaling, no collective on the path), the config-3 training step with the NCCL gradient
# all-reduce.  usage: gpu_multi_r2.sh <tag> <N>
set -u
OUT=gpurun_out; TAG=${1:-r02}; N=${2:-8}; mkdir -p $OUT
nvidia-smi -L | wc -l > $OUT/${TAG}_ngpus.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N"
timeout 400 python -m pytest tests/test_gpu_ddp_nccl.py tests/test_gpu_threads.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_mg_pytest.log 2>&1
tail -2 $OUT/${TAG}_mg_pytest.log
timeout 400 $RUN bench.py --gpus $N --steps 20 --warmup 3 --no-fused --no-gpu-baseline --no-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench N=$N exit $?"; cut -c1-200 $OUT/${TAG}_bench_n$N.json
timeout 600 $RUN scripts/bench_configs.py 5a 5b 3 4 1 > $OUT/${TAG}_configs_n$N.jsonl 2> $OUT/${TAG}_configs_n$N.err
echo "configs N=$N exit $?"; cut -c1-260 $OUT/${TAG}_configs_n$N.jsonl
timeout 400 $RUN scripts/train_step_ddp.py > $OUT/${TAG}_train_ddp_n$N.json 2> $OUT/${TAG}_train_ddp_n$N.err
echo "train N=$N exit $?"; cut -c1-400 $OUT/${TAG}_train_ddp_n$N.json
