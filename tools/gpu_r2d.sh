#!/bin/bash
mkdir -p gpurun_out
MYR_LIB=$PWD/build/lib_prof128.so timeout 200 python tools/phase_profile.py 1024 > gpurun_out/phase.log 2>&1
echo "256x1" >> gpurun_out/phase.log
MYR_LIB=$PWD/build/lib_prof256.so timeout 200 python tools/phase_profile.py 1024 >> gpurun_out/phase.log 2>&1
cat gpurun_out/phase.log
