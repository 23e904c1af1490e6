#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
for v in t128x2 t256x1 t256x2 t512x1; do MYR_LIB=$PWD/build/lib_$v.so timeout 200 python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1; done
echo "t128x2 CTAS=1" >> gpurun_out/ab.log; MYR_IPM_CTAS=1 MYR_LIB=$PWD/build/lib_t128x2.so timeout 200 python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
