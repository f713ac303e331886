#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
tools/gpu_checks.sh tests/test_gpu_obs_transforms.py tests/test_gpu_head_ops.py 2>&1 | grep -v "^PASSED"
grep "resize+crop" gpurun_out/test_gpu_obs_transforms.log
