#!/bin/bash
# Experiment: fp16 raw conv output under GroupNorm (depth trunk traffic).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
for r in 0 1; do
  ( VLNCE_GN_RAW16=$r timeout 300 python tools/perf_probe.py ) > gpurun_out/phases_raw16_$r.log 2>&1; echo "probe raw16=$r exit $?"; tail -1 gpurun_out/phases_raw16_$r.log | cut -c1-400
done
VLNCE_GN_RAW16=1 timeout 300 python -m pytest tests/test_gpu_policy.py tests/test_gpu_train.py -m gpu -q --timeout 300 -p no:cacheprovider -s 2>&1 | grep -E "rel err|passed|failed|Error" | cut -c1-200
( VLNCE_GN_RAW16=1 timeout 300 python bench.py --no-cpu-baseline ) > gpurun_out/bench_raw16.log 2> gpurun_out/bench_raw16.err
echo "bench raw16 exit $?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_raw16.log').read().strip().splitlines()[-1])
print('raw16 value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],4))
PY
