#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/ab.log
timeout 300 python tools/ab_bench.py trap,hs >> gpurun_out/ab.log 2>&1
echo "MYR_IPM_CTAS=1" >> gpurun_out/ab.log; MYR_IPM_CTAS=1 timeout 200 python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1
MYR_LIB=$PWD/build/lib_prof.so timeout 200 python tools/phase_profile.py 1024 > gpurun_out/phase.log 2>&1
echo "CTAS=1" >> gpurun_out/phase.log
MYR_IPM_CTAS=1 MYR_LIB=$PWD/build/lib_prof.so timeout 200 python tools/phase_profile.py 1024 >> gpurun_out/phase.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -f -o gpurun_out/prof_ipm_2cta python tools/profile_run.py trap 296 ipm > gpurun_out/prof_ipm.log 2>&1
MYR_IPM_CTAS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -f -o gpurun_out/prof_ipm_1cta python tools/profile_run.py trap 148 ipm >> gpurun_out/prof_ipm.log 2>&1
cat gpurun_out/ab.log gpurun_out/phase.log; tail -3 gpurun_out/prof_ipm.log
