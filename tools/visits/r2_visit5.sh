#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python tools/trace_depth.py 64 > gpurun_out/trace_depth.log 2>&1; echo "trace exit $?"; head -1 gpurun_out/trace_depth.log; tail -4 gpurun_out/trace_depth.log
timeout 200 python tools/trace_depth.py 1 > gpurun_out/trace_depth1.log 2>&1; echo "trace1 exit $?"; head -1 gpurun_out/trace_depth1.log; tail -4 gpurun_out/trace_depth1.log
