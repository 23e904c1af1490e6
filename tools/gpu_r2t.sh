#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
timeout 300 python tools/ab_bench.py trap,hs 2>&1 | grep -v Warn > gpurun_out/ab.log; cat gpurun_out/ab.log
timeout 300 python tools/quickbench_node.py 2>&1 | grep -v Warn > gpurun_out/quickbench_node.log; tail -2 gpurun_out/quickbench_node.log
for tool in racecheck memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.log; tail -3 gpurun_out/r2_sanitizer_$tool.log
done
