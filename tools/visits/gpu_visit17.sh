#!/bin/bash
# Round-end evidence: full GPU suite, smoke, default bench (+ reference arm), ncu launch list, ncu --set full
# captures of the heavy RGB-trunk conv_tma launches.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
tools/gpu_checks.sh tests/test_gpu_conv.py tests/test_gpu_head_ops.py tests/test_gpu_policy.py tests/test_gpu_train.py tests/test_gpu_waypoint.py 2>&1 | grep -v "^PASSED"
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; grep smoke: gpurun_out/smoke.log
( time timeout 420 python bench.py --profile ) > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench exit $?"; tail -1 gpurun_out/bench_final.log | cut -c1-300; tail -4 gpurun_out/bench_final.err
( time timeout 200 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -4 gpurun_out/bench_ref.log | cut -c1-400
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_step.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 29 -c 8 \
  -o gpurun_out/r01d_conv_tma_rgb -f python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full1.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']); print(d.get('cpu_baseline'))
PY
