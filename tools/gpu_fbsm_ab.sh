#!/bin/bash
# A/B of a sweep-kernel change on one GPU: FBSM GPU tests, quick bench, one ncu full capture (tag = $1)
TAG=${1:-b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fbsm.py -m gpu -q 2>&1 | tail -2
timeout 600 python tools/quickbench_fbsm.py 2>&1 | grep -v Warn > gpurun_out/r2_quickbench_fbsm_$TAG.log; cat gpurun_out/r2_quickbench_fbsm_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbsm_kernel -s 1 -c 1 -f -o gpurun_out/r2_fbsm_cancer_N1000_B65536_$TAG python tools/quickbench_fbsm.py profile CANCERTREATMENT 1000 65536 > gpurun_out/prof_fbsm_$TAG.log 2>&1; tail -1 gpurun_out/prof_fbsm_$TAG.log
