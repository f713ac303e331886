#!/bin/bash
# Waypoint path: kernel-level and end-to-end GPU tests.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
timeout 900 python -m pytest tests/test_gpu_waypoint.py -m gpu -q -rA --timeout 600 -p no:cacheprovider -s > gpurun_out/test_gpu_waypoint.log 2>&1
echo "waypoint tests exit $?"
grep -E "^(PASSED|FAILED|ERROR)|passed|failed|rel err|worst|fused WDDPPO|Error|error:" gpurun_out/test_gpu_waypoint.log | tail -60
