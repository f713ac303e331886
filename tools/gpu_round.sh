#!/bin/bash
# One GPU-box visit: smoke, bench (ours + reference arm), the gpu test files, ncu launch list.
# Everything lands under gpurun_out/.  Usage: tools/gpu_round.sh [skip-ncu]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
( time timeout 420 python bench.py --profile ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
( time timeout 200 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -4 gpurun_out/bench_ref.log
tools/gpu_checks.sh tests/test_gpu_conv.py tests/test_gpu_head_ops.py tests/test_gpu_obs_transforms.py tests/test_gpu_policy.py tests/test_gpu_train.py tests/test_gpu_waypoint.py tests/test_zz_gpu_variants.py
if [ "$1" != "skip-ncu" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"
fi
