#!/bin/bash
# Round 2, visit 1: provable bench (normal exit, parity block, 64-row CPU arm), library baseline (cuDNN/cuBLAS eager,
# 4 arithmetic modes + trunks alone + act), act-only bench, phase probe, N4 tests forced to run.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
( time timeout 600 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['e2e']['value'], json.dumps(d.get('parity')), json.dumps(d.get('cpu_baseline')))
r = d['roofline']; print(r['bound'], r['frac'], r['conv_ms_per_step'], r.get('layer_model', {}).get('frac_of_ideal'))"
tail -3 gpurun_out/bench.err
( time timeout 900 python tests/library_baseline.py ) > gpurun_out/library_baseline.log 2>&1; echo "library baseline exit $?"; tail -1 gpurun_out/library_baseline.log
( time timeout 300 python tools/bench_act.py ) > gpurun_out/bench_act.log 2>&1; echo "bench_act exit $?"; tail -1 gpurun_out/bench_act.log
( time timeout 300 python tools/perf_probe.py ) > gpurun_out/phases.log 2>&1; echo "probe exit $?"; tail -1 gpurun_out/phases.log
timeout 600 python -m pytest tests/test_zz_gpu_variants.py -m gpu -q -rA --runxfail --timeout 300 -p no:cacheprovider -k "trunk" \
  > gpurun_out/test_n4.log 2>&1
echo "n4 exit $?"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed|worst" gpurun_out/test_n4.log | tail -12
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
