#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
( timeout 500 python tools/bench_waypoint.py --envs 4 --steps 16 ) > gpurun_out/bench_waypoint.log 2>&1; echo "wp bench exit $?"; tail -3 gpurun_out/bench_waypoint.log | cut -c1-1500
