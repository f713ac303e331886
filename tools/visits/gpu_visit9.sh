#!/bin/bash
# Evidence refresh for the conv_tma v3 kernel mix: launch list of one eager step + two ncu --set full
# captures of the dominant RGB-trunk tiles (3x3 and 1x1+residual).  Outputs under gpurun_out/.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_step.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu launches exit $?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 3 -c 6 \
  -o gpurun_out/r01c_conv_tma_step -f python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full1.log
ls -la gpurun_out | grep r01c
