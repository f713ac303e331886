#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "tma or s2d" > gpurun_out/t_conv_new.log 2>&1
echo "new conv tests exit $?"; tail -15 gpurun_out/t_conv_new.log
( timeout 400 python tools/conv_layers.py --reps 10 ) > gpurun_out/conv_sweep.log 2>&1; echo "sweep exit $?"; tail -40 gpurun_out/conv_sweep.log
( timeout 300 python tools/perf_probe.py ) > gpurun_out/phases.log 2>&1; echo "probe exit $?"; tail -3 gpurun_out/phases.log
( timeout 300 python bench.py --profile --no-cpu-baseline ) > gpurun_out/bench3.log 2> gpurun_out/bench3.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench3.log | tail -2; tail -3 gpurun_out/bench3.err
tools/gpu_checks.sh tests/test_gpu_policy.py tests/test_gpu_train.py
