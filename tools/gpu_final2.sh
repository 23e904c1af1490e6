#!/bin/bash
# final check of the round-2 tree on one GPU: whole GPU suite, smoke, bench line, FBSM quick bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -2 gpurun_out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; cut -c1-300 gpurun_out/r2_bench_1gpu_final.json; tail -2 gpurun_out/r2_bench_1gpu_final.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
timeout 600 python tools/quickbench_fbsm.py 2>&1 | grep -v Warn > gpurun_out/r2_quickbench_fbsm.log; tail -4 gpurun_out/r2_quickbench_fbsm.log
