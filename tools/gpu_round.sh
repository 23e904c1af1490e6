#!/bin/bash
# One gpurun call: GPU parity tests, quick benches, headline bench, ncu launch list and full captures.
set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 300 python tools/quickbench.py > gpurun_out/quickbench.log 2>&1
timeout 300 python tools/quickbench_shooting.py > gpurun_out/quickbench_shooting.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 2 -o gpurun_out/prof_k1 python tools/profile_run.py trap 8192 k1 > gpurun_out/prof_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -o gpurun_out/prof_ipm python tools/profile_run.py trap 1024 ipm > gpurun_out/prof_ipm.log 2>&1
ls -la gpurun_out
