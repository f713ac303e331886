#!/bin/bash
# 2-GPU visit: NCCL tests (IL + DD-PPO steps), bench with the all-reduce captured in the graph vs eager, C4 loop, C5 workload
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export VLNCE_AUTOTUNE_CACHE=/tmp/vlnce_tune.json
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q --timeout 500 -p no:cacheprovider 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 2> gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.log; echo "bench n2 exit $?"
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --eager-allreduce 2> gpurun_out/bench_n2_eager.err | tail -1 > gpurun_out/bench_n2_eager.log; echo "bench n2 eager exit $?"
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-parity 2> gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.log
for f in n1 n2 n2_eager; do python -c "
import json
d=json.loads(open('gpurun_out/bench_$f.log').read()); print('$f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['config'].get('allreduce_in_graph'), d['config']['cuda_graph'])"; done
tail -2 gpurun_out/bench_n2.err
timeout 400 $TR tools/bench_c4.py --iters 3 > gpurun_out/bench_c4_n2.log 2>&1; echo "c4 n2 exit $?"; tail -1 gpurun_out/bench_c4_n2.log | cut -c1-400
timeout 300 python tools/bench_c4.py --iters 3 > gpurun_out/bench_c4_n1.log 2>&1; echo "c4 n1 exit $?"; tail -1 gpurun_out/bench_c4_n1.log | cut -c1-400
timeout 400 python bench.py --workload c5 --steps 20 --warmup 4 2> gpurun_out/bench_c5_n1.err | tail -1 > gpurun_out/bench_c5_n1.log; echo "c5 exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_c5_n1.log').read()); print('c5 n1', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), json.dumps(d.get('parity'))[:300])"
tail -2 gpurun_out/bench_c5_n1.err
