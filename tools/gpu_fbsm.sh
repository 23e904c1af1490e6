#!/bin/bash
# FBSM evidence on one GPU: GPU test suite, smoke, sweep-kernel bench + roofline, ncu full capture + launch list, sanitizer
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python tools/quickbench_fbsm.py 2>&1 | grep -v Warn > gpurun_out/r2_quickbench_fbsm.log; cat gpurun_out/r2_quickbench_fbsm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbsm_kernel -s 1 -c 1 -f -o gpurun_out/r2_fbsm_cancer_N1000_B65536 python tools/quickbench_fbsm.py profile CANCERTREATMENT 1000 65536 > gpurun_out/prof_fbsm.log 2>&1; tail -2 gpurun_out/prof_fbsm.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/quickbench_fbsm.py profile PREDATORPREY 100 2000 > gpurun_out/r2_sanitizer_fbsm_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_fbsm_memcheck.log; tail -3 gpurun_out/r2_sanitizer_fbsm_memcheck.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; cut -c1-600 gpurun_out/r2_bench_1gpu_final.json; tail -2 gpurun_out/r2_bench_1gpu_final.err
