#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_depth_fused.py -m gpu -x -q -s --timeout 180 -p no:cacheprovider 2>&1 | grep -E "rel err|fused vs|passed|failed|Error|vlnce:" | head
timeout 200 python tools/trace_depth.py 64 > gpurun_out/trace_depth.log 2>&1; echo "trace exit $?"; head -1 gpurun_out/trace_depth.log
( time timeout 300 python tools/perf_probe.py ) > gpurun_out/phases.log 2>&1; echo "probe exit $?"; tail -1 gpurun_out/phases.log
