#!/bin/bash
# full GPU validation of the round-2 tree + evidence (post racecheck fix)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python tools/ab_bench.py trap,hs > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log
timeout 300 python tools/quickbench_shooting.py > gpurun_out/quickbench_shooting.log 2>&1; cat gpurun_out/quickbench_shooting.log
timeout 300 python tools/quickbench_node.py > gpurun_out/quickbench_node.log 2>&1; tail -4 gpurun_out/quickbench_node.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for tool in racecheck memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.log; tail -3 gpurun_out/r2_sanitizer_$tool.log
done
