#!/bin/bash
# round-2 end-state visit (second pass): full GPU suite, smoke, bench (ours / reference arm), per-kernel ncu table
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_suite.log 2>&1; echo "suite rc $?"; tail -2 gpurun_out/final_gpu_suite.log
python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/final_bench.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc $?"; tail -1 gpurun_out/final_bench_ref.log | cut -c1-160
M=$(python -c "import sys; sys.path.insert(0,'tools'); import summarize_profiles as s; print(s.KERNEL_METRICS)")
timeout 600 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/kernels_step.csv --metrics $M python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_kernels.log 2>&1; echo "ncu rc $?"
