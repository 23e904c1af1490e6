#!/bin/bash
mkdir -p gpurun_out
for v in "" build/lib_th1.so build/lib_th13.so build/lib_th50.so build/lib_th100.so; do
  MYR_LIB=$v timeout 300 python tools/ab_bench.py trap,hs 2>&1 | grep -v Warn
done > gpurun_out/ab_thomas.log 2>&1
cat gpurun_out/ab_thomas.log
MYR_LIB=build/lib_th26ph.so timeout 300 python tools/phase_profile.py 1024 > gpurun_out/phase_th26.log 2>&1; cat gpurun_out/phase_th26.log
