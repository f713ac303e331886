#!/bin/bash
# 2-GPU visit: NCCL gradient all-reduce test + the torchrun bench at N=2 (weak scaling check).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rA --timeout 300 -p no:cacheprovider > gpurun_out/test_gpu_dist.log 2>&1
echo "dist test exit $?"; tail -5 gpurun_out/test_gpu_dist.log | cut -c1-300
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "bench N=2 exit $?"; tail -1 gpurun_out/bench_n2.log | cut -c1-400; tail -3 gpurun_out/bench_n2.err | cut -c1-300
