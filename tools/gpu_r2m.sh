#!/bin/bash
mkdir -p gpurun_out
python tools/seir_debug.py x_seir_trap_10 > gpurun_out/seir_default.log 2>&1; tail -2 gpurun_out/seir_default.log
MYR_LIB=build/lib_seirtrace.so python tools/seir_debug.py x_seir_trap_10 > gpurun_out/seir_trace.log 2>&1; tail -3 gpurun_out/seir_trace.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -15 gpurun_out/pytest_gpu_all.log
