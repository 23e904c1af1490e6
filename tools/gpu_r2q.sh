#!/bin/bash
mkdir -p gpurun_out
{
echo "1 CTA/SM:"; MYR_IPM_CTAS=1 MYR_LIB=build/lib_base.so timeout 300 python tools/ab_bench.py trap 2>&1 | grep -v Warn
echo "1 CTA/SM, 115 KB smem cap:"; MYR_IPM_SMEM_KB=113 MYR_IPM_CTAS=1 MYR_LIB=build/lib_base.so timeout 300 python tools/ab_bench.py trap 2>&1 | grep -v Warn
} > gpurun_out/ab_occ1.log 2>&1
cat gpurun_out/ab_occ1.log
