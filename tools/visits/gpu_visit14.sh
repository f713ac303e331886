#!/bin/bash
# Experiment: depth trunk as independent batch chunks on parallel streams.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
( timeout 500 python tools/perf_probe.py --chunks 1,2,4,8 ) > gpurun_out/phases_chunks.log 2>&1; echo "probe exit $?"; grep -E "^chunks|rgb_trunk|Error|error" gpurun_out/phases_chunks.log | cut -c1-300
VLNCE_DEPTH_CHUNKS=4 timeout 300 python -m pytest tests/test_gpu_policy.py -m gpu -q --timeout 300 -p no:cacheprovider -k "trunks or cma_act" 2>&1 | tail -3
for k in 2 4; do
  ( VLNCE_DEPTH_CHUNKS=$k timeout 300 python bench.py --no-cpu-baseline ) > gpurun_out/bench_chunks$k.log 2> gpurun_out/bench_chunks$k.err
  echo "bench chunks=$k exit $?"; tail -2 gpurun_out/bench_chunks$k.err | cut -c1-200; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_chunks$k.log').read().strip().splitlines()[-1])
print($k, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
PY
done
