#!/bin/bash
# Runs the GPU test files in separate processes (a trapped kernel only poisons its own
# process) and keeps the logs under gpurun_out/.  Usage: tools/gpu_checks.sh [files...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/gpu.txt 2>&1
FILES="$@"
[ -z "$FILES" ] && FILES="tests/test_gpu_conv.py tests/test_gpu_head_ops.py tests/test_gpu_obs_transforms.py tests/test_gpu_policy.py tests/test_gpu_train.py tests/test_gpu_waypoint.py tests/test_zz_gpu_variants.py"
rc=0
for f in $FILES; do
  name=$(basename $f .py)
  extra=""
  [ "$name" = "test_gpu_conv" ] && extra="-n 4"
  timeout 900 python -m pytest $f -m gpu -q -rA --timeout 300 $extra -p no:cacheprovider > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f exit $r"
  grep -E "^(PASSED|FAILED|ERROR)|passed|failed|error|rel err|vlnce:" gpurun_out/$name.log | tail -40
  [ $r -ne 0 ] && rc=$r
done
exit $rc
