#!/bin/bash
# Experiment: constant max-shared carve-out for the chained kernels (no SM reconfiguration between kernels).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
for r in 0 1; do
  ( VLNCE_CARVEOUT=$r timeout 300 python tools/perf_probe.py ) > gpurun_out/phases_carve_$r.log 2>&1; echo "probe carveout=$r exit $?"; tail -1 gpurun_out/phases_carve_$r.log | cut -c1-400
done
