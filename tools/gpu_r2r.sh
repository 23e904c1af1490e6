#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/ab_bench.py trap,hs 2>&1 | grep -v Warn
timeout 300 python tools/quickbench_shooting.py 2>&1 | grep -v Warn
} > gpurun_out/ab_layout.log 2>&1
cat gpurun_out/ab_layout.log
