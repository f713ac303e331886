#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 120 python tools/trace_conv.py l3c1 ) > gpurun_out/trace_l3c1.log 2>&1; echo "exit $?"
( TRACE_LINES=200 timeout 120 python tools/trace_conv.py stem "64|(4<<16)|(1<<21)" ) > gpurun_out/trace_stem.log 2>&1; echo "exit $?"
( TRACE_LINES=200 timeout 120 python tools/trace_conv.py l2c2 "128|(4<<16)|(1<<21)" ) > gpurun_out/trace_l2c2.log 2>&1; echo "exit $?"
head -150 gpurun_out/trace_l3c1.log
