M=$(python -c "import sys; sys.path.insert(0,'tools'); import summarize_profiles as s; print(s.KERNEL_METRICS)")
timeout 600 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/kernels_step.csv --metrics $M python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_kernels.log 2>&1; echo "ncu rc $?"
timeout 400 python bench.py --workload c2t8 --no-cpu-baseline > gpurun_out/bench_c2t8.log 2>&1; echo "c2t8 rc $?"; tail -1 gpurun_out/bench_c2t8.log | cut -c1-300
timeout 400 python bench.py --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.log 2>&1; echo "tf32 rc $?"; tail -1 gpurun_out/bench_tf32.log | cut -c1-300
