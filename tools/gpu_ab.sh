#!/bin/bash
# tools/gpu_ab.sh <tag> [quads]: A/B bench of build/lib_<tag>.so (CARTPOLE-only variant)
mkdir -p gpurun_out
MYR_LIB=$PWD/build/lib_$1.so timeout 300 python tools/ab_bench.py ${2:-trap} 2>&1 | grep -v Warn | tee gpurun_out/ab_$1.log
