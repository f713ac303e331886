#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_depth_fused.py -m gpu -x -q --timeout 180 -p no:cacheprovider 2>&1 | tail -2
timeout 200 python tools/trace_depth.py 64 > gpurun_out/trace_depth.log 2>&1; echo "trace exit $?"; head -1 gpurun_out/trace_depth.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:depth_fused -s 3 -c 1 -f -o gpurun_out/r02_depth_fused python tools/trace_depth.py 64 > gpurun_out/ncu_depth.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_depth.log
