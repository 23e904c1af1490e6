#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do MYR_LIB=$PWD/build/lib_$v.so timeout 300 python tools/ab_bench.py trap,hs 2>&1 | grep -v Warn | grep "B=1024\|B=8192"; done | tee gpurun_out/ab3.log
