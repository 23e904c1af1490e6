#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
timeout 300 python tools/ab_bench.py trap,hs 2>&1 | grep -v Warn > gpurun_out/ab.log; cat gpurun_out/ab.log
timeout 300 python tools/quickbench_shooting.py 2>&1 | grep -v Warn > gpurun_out/quickbench_shooting.log; cat gpurun_out/quickbench_shooting.log
timeout 300 python tools/quickbench_node.py 2>&1 | grep -v Warn > gpurun_out/quickbench_node.log; tail -2 gpurun_out/quickbench_node.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-900 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
