#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
timeout 600 python -m pytest tests/test_gpu_waypoint.py -m gpu -x -q --timeout 500 -p no:cacheprovider -k "fused_wddppo or ppo_minibatch" 2>&1 | tail -3
timeout 300 python tools/bench_c4.py --iters 3 > gpurun_out/bench_c4_n1.log 2>&1; echo "c4 n1 exit $?"; tail -1 gpurun_out/bench_c4_n1.log | cut -c1-330
VLNCE_PPO_GRAPH=0 timeout 300 python tools/bench_c4.py --iters 3 > gpurun_out/bench_c4_n1_eager.log 2>&1; tail -1 gpurun_out/bench_c4_n1_eager.log | cut -c1-330
timeout 900 python tools/bench_c3.py --parity > gpurun_out/bench_c3.log 2>&1; echo "c3 exit $?"; tail -2 gpurun_out/bench_c3.log | cut -c1-1200
