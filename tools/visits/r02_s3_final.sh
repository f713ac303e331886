#!/bin/bash
# round-2 end-state visit: full GPU suite, smoke, bench (ours / reference arm / tf32), launch list with DRAM bytes, sanitizer
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_suite.log 2>&1; echo "suite rc $?"; tail -2 gpurun_out/final_gpu_suite.log
python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/final_bench.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc $?"; tail -1 gpurun_out/final_bench_ref.log | cut -c1-200
python bench.py --precision tf32 --no-cpu-baseline > gpurun_out/final_bench_tf32.log 2>&1; echo "tf32 rc $?"; tail -1 gpurun_out/final_bench_tf32.log | cut -c1-200
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/final_launches_step.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1; echo "ncu rc $?"
bash tools/sanitize.sh 2>&1 | tail -8
