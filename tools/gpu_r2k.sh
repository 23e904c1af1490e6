#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
for v in t128x2 t256x1; do MYR_LIB=$PWD/build/lib_$v.so timeout 200 python tools/ab_bench.py trap,hs >> gpurun_out/ab.log 2>&1; done
MYR_LIB=$PWD/build/lib_prof128.so timeout 200 python tools/phase_profile.py 1024 > gpurun_out/phase.log 2>&1
cat gpurun_out/ab.log gpurun_out/phase.log
