#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "tma or s2d" > gpurun_out/t_conv_new.log 2>&1
echo "new conv tests exit $?"; tail -3 gpurun_out/t_conv_new.log
( TRACE_LINES=120 timeout 120 python tools/trace_conv.py l3c1 ) > gpurun_out/trace_l3c1.log 2>&1; echo "exit $?"
( TRACE_LINES=200 timeout 120 python tools/trace_conv.py stem "64|(4<<16)|(1<<21)" ) > gpurun_out/trace_stem.log 2>&1; echo "exit $?"
( TRACE_LINES=200 timeout 120 python tools/trace_conv.py l2c2 "128|(4<<16)|(1<<21)" ) > gpurun_out/trace_l2c2.log 2>&1; echo "exit $?"
( timeout 400 python tools/conv_layers.py --reps 10 --only "stem,l3.c1,l2.c2 3x3 128,l1.c3,l3.c2 3x3 256,l1.c1 1x1 256,l4.c3,l2.c1" ) > gpurun_out/conv_sweep.log 2>&1; echo "sweep exit $?"
sed -E 's/bn(32|64|128|256):[^b]*//g' gpurun_out/conv_sweep.log | tail -12
sed -n 60,100p gpurun_out/trace_l3c1.log
