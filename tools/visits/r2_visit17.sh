#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_conv.py -m gpu -x -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
( time timeout 300 python tools/perf_probe.py ) > gpurun_out/phases.log 2>&1; echo "probe exit $?"; tail -1 gpurun_out/phases.log
