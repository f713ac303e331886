#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
timeout 200 python -m pytest tests/test_gpu_depth_fused.py -m gpu -x -q --timeout 180 -p no:cacheprovider 2>&1 | tail -1
for cap in 0 140 132 120; do
  VLNCE_TMA_CTA_CAP=$cap timeout 300 python bench.py --no-cpu-baseline --no-parity --steps 30 2> gpurun_out/bench_cap$cap.err | tail -1 > gpurun_out/bench_cap$cap.log
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_cap$cap.log').read())
print('cap $cap: value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
done
