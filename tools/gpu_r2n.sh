#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -5 gpurun_out/pytest_gpu_all.log
timeout 300 python tools/ab_bench.py trap,hs > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log
