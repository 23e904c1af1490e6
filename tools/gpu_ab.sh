#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab.log
python tools/ab_bench.py trap,hs >> gpurun_out/ab.log 2>&1
for v in "$@"; do MYR_LIB=$PWD/build/lib_$v.so python tools/ab_bench.py trap >> gpurun_out/ab.log 2>&1; done
cat gpurun_out/ab.log
