#!/bin/bash
# Diagnostics visit: phase probe, per-kernel ncu sections over ONE eager step, launch list, full captures
# of selected conv kernels.  Outputs (text/CSV + a few .ncu-rep) under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
( time timeout 300 python tools/perf_probe.py ) > gpurun_out/phases.log 2>&1; echo "probe exit $?"; tail -4 gpurun_out/phases.log
( time timeout 300 python bench.py --profile --no-cpu-baseline ) > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/bench2.log | tail -2
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_step.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --profile-from-start off --clock-control none -k regex:conv \
  --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --section WarpStateStats --section SchedulerStats --csv --page raw \
  --log-file gpurun_out/conv_sections.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_sections.log 2>&1
echo "ncu sections exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 5 -c 1 \
  -o gpurun_out/r01b_conv_tma_3x3_l2 python tools/conv_layers.py --only "l2.c2 3x3 128" --reps 2 > gpurun_out/ncu_full1.log 2>&1
echo "ncu full1 exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 5 -c 1 \
  -o gpurun_out/r01b_conv_tma_1x1res_l2 python tools/conv_layers.py --only "l2.c3" --reps 2 > gpurun_out/ncu_full2.log 2>&1
echo "ncu full2 exit $?"
ls -la gpurun_out | head -40
