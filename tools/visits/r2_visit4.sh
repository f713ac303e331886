#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_depth_fused.py -m gpu -x -q -rA -s --timeout 240 -p no:cacheprovider > gpurun_out/test_depth_fused.log 2>&1
echo "depth_fused tests exit $?"; grep -E "rel err|fused vs|FAILED|ERROR|passed|failed|vlnce:|Error|error" gpurun_out/test_depth_fused.log | head -20
timeout 200 python tools/trace_depth.py 64 > gpurun_out/trace_depth.log 2>&1; echo "trace exit $?"; head -3 gpurun_out/trace_depth.log
( time timeout 300 python tools/perf_probe.py ) > gpurun_out/phases.log 2>&1; echo "probe exit $?"; tail -1 gpurun_out/phases.log
