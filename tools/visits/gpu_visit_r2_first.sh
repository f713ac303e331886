#!/bin/bash
# First GPU visit of round 2: everything written after round 1's GPU budget ran out.
#   1. the late-collected variant tests with the non-strict xfails forced to run as normal tests (--runxfail):
#      trunk-backward kernels + trainable-trunk goldens (N4), LSTM / GRU / normalize_rgb end-to-end goldens;
#   2. the default bench line — its `roofline` object now comes from the per-launch model (bound "hbm");
#   3. the library baseline (oracle torch modules = cuDNN / cuBLAS in CUDA eager mode: fp32 / tf32 / bf16 autocast);
#   4. the usual suite.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_variants.py -m gpu -q -rA --runxfail --timeout 300 -p no:cacheprovider \
  > gpurun_out/test_zz_gpu_variants.log 2>&1
echo "variants exit $?"
grep -E "^(PASSED|FAILED|ERROR)|passed|failed|trunk:|worst" gpurun_out/test_zz_gpu_variants.log | tail -40
( time timeout 420 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], json.dumps(d['roofline'])[:900])"
( time timeout 300 python tools/bench_act.py ) > gpurun_out/bench_act.log 2>&1; echo "bench_act exit $?"; tail -1 gpurun_out/bench_act.log
( time timeout 600 python tests/library_baseline.py ) > gpurun_out/library_baseline.log 2>&1; echo "library baseline exit $?"; tail -2 gpurun_out/library_baseline.log
tools/gpu_checks.sh 2>&1 | grep -v "^PASSED"
# refresh the launch list of one serial step and the per-layer gap table derived from it
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_step.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu exit $?"
python tools/roofline_model.py --launch-csv gpurun_out/launches_step.csv --gap-md gpurun_out/layer_gap.md > gpurun_out/roofline_model.log 2>&1
tail -6 gpurun_out/layer_gap.md
