#!/bin/bash
# Experiment: leave a few SMs free of the persistent RGB conv CTAs so the depth trunk runs beside them.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
( timeout 400 python tools/perf_probe.py --caps 0,140,132,124,116,104 ) > gpurun_out/phases_caps.log 2>&1; echo "probe exit $?"; grep -E "^cap|rgb_trunk" gpurun_out/phases_caps.log | cut -c1-400
for cap in 132 120; do
  ( VLNCE_TMA_CTA_CAP=$cap timeout 300 python bench.py --no-cpu-baseline ) > gpurun_out/bench_cap$cap.log 2> gpurun_out/bench_cap$cap.err
  echo "bench cap=$cap exit $?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cap$cap.log').read().strip().splitlines()[-1])
print($cap, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
PY
done
