#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/ab_bench.py trap,hs > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log
timeout 300 python tools/quickbench_shooting.py > gpurun_out/quickbench_shooting.log 2>&1; cat gpurun_out/quickbench_shooting.log
timeout 300 python tools/quickbench_node.py > gpurun_out/quickbench_node.log 2>&1; tail -4 gpurun_out/quickbench_node.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
