#!/bin/bash
# compute-sanitizer passes over the tcgen05 / TMA conv tiles and the head kernels (run on a B200; logs -> gpurun_out/,
# summaries committed under profiles/).  The tools slow kernels 10-100x: bf16 cases only.
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "(tma_persistent and bf16 and (auto or mt2)) or (s2_tma and bf16 and auto) or (halo and bf16) or (s2d_stem and bf16 and auto and 40)" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck conv rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "(tma_persistent and bf16 and auto) or (s2_tma and bf16 and auto)" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck conv rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_head_ops.py -m gpu -q -x -k "sgemm or lstm or small_layout" > gpurun_out/sanitizer_memcheck_head.log 2>&1; echo "memcheck head rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_head.log
