#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
for v in "$@"; do MYR_LIB=$PWD/build/lib_$v.so timeout 200 python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1; done
cat gpurun_out/ab.log
