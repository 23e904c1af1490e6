#!/bin/bash
mkdir -p gpurun_out
{
for t in 8 0 4 6 12; do echo "MYR_IPM_TURN=$t"; MYR_IPM_TURN=$t MYR_LIB=$PWD/build/lib_$1.so timeout 300 python tools/ab_bench.py ${2:-trap} 2>&1 | grep -v Warn; done
} | tee gpurun_out/ab_turn_$1.log
