#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_waypoint.py tests/test_gpu_train.py tests/test_zz_gpu_variants.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/parity_suite.log 2>&1; echo "exit $?"
grep -E "rel err|rel-L2|worst|passed|failed|FAILED|Error" gpurun_out/parity_suite.log | head -80
