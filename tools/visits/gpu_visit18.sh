#!/bin/bash
# waypoint GPU tests (after the test fix) + ncu --set full captures of RGB-trunk conv tiles in situ.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
tools/gpu_checks.sh tests/test_gpu_waypoint.py 2>&1 | grep -v "^PASSED"
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_ -s 62 -c 10 \
  -o gpurun_out/r01d_conv_rgb -f python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full1.log; ls -la gpurun_out | grep r01d
