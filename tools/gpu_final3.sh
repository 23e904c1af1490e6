#!/bin/bash
# last check of the round-2 tree on one GPU: whole GPU suite, smoke, bench line, memcheck of every kernel at small sizes
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -2 gpurun_out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; cut -c1-300 gpurun_out/r2_bench_1gpu_final.json; tail -2 gpurun_out/r2_bench_1gpu_final.err
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_memcheck_final.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck_final.log; tail -8 gpurun_out/r2_sanitizer_memcheck_final.log
