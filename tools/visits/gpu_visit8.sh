#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
tools/gpu_checks.sh tests/test_gpu_train.py tests/test_gpu_policy.py tests/test_gpu_conv.py tests/test_gpu_head_ops.py 2>&1 | grep -v "^PASSED"
( timeout 300 python tools/perf_probe.py ) > gpurun_out/phases.log 2>&1; echo "probe exit $?"; tail -2 gpurun_out/phases.log
( timeout 300 python bench.py --profile --no-cpu-baseline ) > gpurun_out/bench3.log 2> gpurun_out/bench3.err; echo "bench exit $?"; cut -c1-330 gpurun_out/bench3.log | tail -1; tail -3 gpurun_out/bench3.err
python - <<'PY'
import json
l=open('gpurun_out/bench3.log').read().strip().splitlines()[-1]
d=json.loads(l); print(d['roofline']); print(d['e2e'])
PY
