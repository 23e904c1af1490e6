#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "seir or SEIR or node or NODE or twin" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/quickbench_node.py > gpurun_out/quickbench_node.log 2>&1; tail -4 gpurun_out/quickbench_node.log
