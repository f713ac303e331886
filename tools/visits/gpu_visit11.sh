#!/bin/bash
# Full GPU suite (default launch mode), then the conv/policy/train/waypoint suites and the bench with
# programmatic dependent launch on, for an A/B.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
tools/gpu_checks.sh tests/test_gpu_conv.py tests/test_gpu_head_ops.py tests/test_gpu_policy.py tests/test_gpu_train.py tests/test_gpu_waypoint.py 2>&1 | grep -v "^PASSED"
echo "=== PDL on"
for f in test_gpu_conv test_gpu_policy test_gpu_train test_gpu_waypoint; do
  extra=""; [ "$f" = "test_gpu_conv" ] && extra="-n 4"
  VLNCE_PDL=1 timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 300 $extra -p no:cacheprovider > gpurun_out/${f}_pdl.log 2>&1
  echo "== $f (PDL) exit $?"; tail -3 gpurun_out/${f}_pdl.log | cut -c1-300
done
for pdl in 0 1; do
  ( timeout 300 python bench.py --pdl $pdl --no-cpu-baseline ) > gpurun_out/bench_pdl$pdl.log 2> gpurun_out/bench_pdl$pdl.err
  echo "bench pdl=$pdl exit $?"; cut -c1-200 gpurun_out/bench_pdl$pdl.log | tail -1; tail -2 gpurun_out/bench_pdl$pdl.err
done
python - <<'PY'
import json
for pdl in (0,1):
    try:
        d=json.loads(open(f'gpurun_out/bench_pdl{pdl}.log').read().strip().splitlines()[-1])
        print(pdl, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],4), 'conv_ms', round(d['roofline']['conv_ms_per_step'],3), d['config'].get('pdl'))
    except Exception as e:
        print(pdl, 'ERR', e)
PY
