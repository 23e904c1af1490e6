#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/prof_k1 python tools/profile_run.py trap 8192 k1 > gpurun_out/prof_k1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -o gpurun_out/prof_ipm python tools/profile_run.py trap 1024 ipm > gpurun_out/prof_ipm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -o gpurun_out/prof_node_k1 python tools/profile_node.py 592 > gpurun_out/prof_node_k1.log 2>&1
ls -la gpurun_out | head -30
