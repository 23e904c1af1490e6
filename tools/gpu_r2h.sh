#!/bin/bash
# round-2 evidence: ncu full captures + launch list + compute-sanitizer logs
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -f -o gpurun_out/r2_ipm_trap_B1024 python tools/profile_run.py trap 1024 ipm > gpurun_out/prof1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -f -o gpurun_out/r2_ipm_hs_B148 python tools/profile_run.py hs 148 ipm > gpurun_out/prof2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for tool in racecheck memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.log; tail -5 gpurun_out/r2_sanitizer_$tool.log
done
ls -la gpurun_out | tail -12
