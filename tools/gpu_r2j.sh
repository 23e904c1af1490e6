#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
MYR_LIB=$PWD/build/lib_t128x2.so timeout 200 python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1
MYR_LIB=$PWD/build/lib_prof128.so timeout 200 python tools/phase_profile.py 1024 > gpurun_out/phase.log 2>&1
cat gpurun_out/ab.log gpurun_out/phase.log
