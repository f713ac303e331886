#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L="stem,l3.c1 1x1 1024,l2.c2 3x3 128,l1.c3,l3.c2 3x3 256,l1.c1 1x1 256"
run() { echo "=== $1"; shift; ( env "$@" timeout 200 python tools/conv_layers.py --reps 10 --only "$L" 2>&1 | grep -v "^sum" | sed -E 's/bn(32|64|128|256):[^b]*//g' ) ; }
{
run default X=1
run deepA6B6 VLNCE_TMA_AS=6 VLNCE_TMA_BS=6 VLNCE_TMA_STG=1
run deepA8B8 VLNCE_TMA_AS=8 VLNCE_TMA_BS=8 VLNCE_TMA_STG=1
run ctas2 VLNCE_TMA_CTAS=2 VLNCE_TMA_AS=2 VLNCE_TMA_BS=3 VLNCE_TMA_STG=1
run ctas2deep VLNCE_TMA_CTAS=2 VLNCE_TMA_AS=3 VLNCE_TMA_BS=4 VLNCE_TMA_STG=1
} > gpurun_out/ring_sweep.log 2>&1
cat gpurun_out/ring_sweep.log
