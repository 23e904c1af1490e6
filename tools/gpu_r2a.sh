#!/bin/bash
# round 2, first GPU call of engine v2: parity tests, A/B of launch shapes, phase profile, other configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/ab.log
timeout 300 python tools/ab_bench.py trap,hs >> gpurun_out/ab.log 2>&1
for v in t192 t256; do MYR_LIB=$PWD/build/lib_$v.so timeout 200 python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1; done
for k in 1 3; do echo "MYR_IPM_CTAS=$k" >> gpurun_out/ab.log; MYR_IPM_CTAS=$k timeout 200 python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1; done
echo "t256 MYR_IPM_CTAS=1" >> gpurun_out/ab.log; MYR_IPM_CTAS=1 MYR_LIB=$PWD/build/lib_t256.so timeout 200 python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1
MYR_LIB=$PWD/build/lib_prof.so timeout 200 python tools/phase_profile.py 1024 > gpurun_out/phase.log 2>&1
timeout 300 python tools/quickbench_shooting.py > gpurun_out/quickbench_shooting.log 2>&1
timeout 300 python tools/quickbench_node.py > gpurun_out/quickbench_node.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/ab.log gpurun_out/phase.log gpurun_out/quickbench_shooting.log gpurun_out/quickbench_node.log
