#!/bin/bash
mkdir -p gpurun_out
{
for v in base mb3 t96mb3 t64mb4; do
  MYR_LIB=build/lib_$v.so timeout 300 python tools/ab_bench.py trap 2>&1 | grep -v Warn
done
echo "forced 4 CTAs/SM:"; MYR_IPM_CTAS=4 MYR_LIB=build/lib_t64mb4.so timeout 300 python tools/ab_bench.py trap 2>&1 | grep -v Warn
echo "forced 3 CTAs/SM:"; MYR_IPM_CTAS=3 MYR_LIB=build/lib_t64mb4.so timeout 300 python tools/ab_bench.py trap 2>&1 | grep -v Warn
} > gpurun_out/ab_occ.log 2>&1
cat gpurun_out/ab_occ.log
