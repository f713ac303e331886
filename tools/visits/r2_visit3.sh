#!/bin/bash
# Round 2, visit 3: first run of the image-stationary depth trunk (csrc/depth_fused.cu) + phase probe.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_depth_fused.py -m gpu -x -q -rA -s --timeout 240 -p no:cacheprovider > gpurun_out/test_depth_fused.log 2>&1
echo "depth_fused tests exit $?"; grep -E "rel err|fused vs|PASSED|FAILED|ERROR|passed|failed|vlnce:|Error|error" gpurun_out/test_depth_fused.log | head -40
( time timeout 300 python tools/perf_probe.py ) > gpurun_out/phases.log 2>&1; echo "probe exit $?"; tail -1 gpurun_out/phases.log
