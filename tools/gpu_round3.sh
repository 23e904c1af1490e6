#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 bash tools/gpu_ab.sh > /dev/null 2>&1
timeout 200 python tools/quickbench_shooting.py > gpurun_out/quickbench_shooting.log 2>&1
timeout 200 python tools/quickbench_node.py trap 1024,4096 > gpurun_out/quickbench_node.log 2>&1
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
