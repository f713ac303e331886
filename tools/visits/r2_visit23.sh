#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
VLNCE_TILE_TABLE=0 timeout 900 python tools/make_tile_table.py gpurun_out/tile_table_b200.json 2>&1 | tail -4
cp gpurun_out/tile_table_b200.json vlnce_b200/tile_table_b200.json
timeout 900 python tools/bench_c3.py > gpurun_out/bench_c3.log 2>&1; echo "c3 exit $?"; tail -1 gpurun_out/bench_c3.log | cut -c1-500
timeout 300 python bench.py --no-cpu-baseline --no-parity --steps 30 2> gpurun_out/bench_tt.err | tail -1 > gpurun_out/bench_tt.log
python -c "
import json
d=json.loads(open('gpurun_out/bench_tt.log').read()); print('bench with table', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
( time timeout 200 python __graft_entry__.py --smoke ) 2>&1 | tail -5
