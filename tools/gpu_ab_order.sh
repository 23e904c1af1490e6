#!/bin/bash
mkdir -p gpurun_out
{
for off in 0 1024 3072 4096; do
  for ord in 0 1; do echo "ROW_OFFSET=$off MYR_ORDER=$ord"; ROW_OFFSET=$off MYR_ORDER=$ord MYR_LIB=$PWD/build/lib_ord.so timeout 300 python tools/ab_bench.py trap 2>&1 | grep "B=1024"; done
done
} | tee gpurun_out/ab_order.log
