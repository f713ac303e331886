#!/bin/bash
# full GPU suite + default bench with the fused depth trunk as the default path
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/gpu_suite.log 2>&1; echo "suite exit $?"; tail -4 gpurun_out/gpu_suite.log
( time timeout 600 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], json.dumps(d.get('parity'))[:400])
r = d['roofline']; print(r['bound'], r['frac'], r['conv_ms_per_step'], r.get('layer_model', {}).get('frac_of_ideal'))"
tail -3 gpurun_out/bench.err
