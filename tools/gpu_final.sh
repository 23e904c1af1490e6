#!/bin/bash
# round-2 final evidence on one GPU: tests, bench line, quick benches, phase cycles, ncu captures, launch list, sanitizers
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; tail -2 gpurun_out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; cut -c1-400 gpurun_out/r2_bench_1gpu.json; tail -2 gpurun_out/r2_bench_1gpu.err
timeout 300 python tools/ab_bench.py trap,hs 2>&1 | grep -v Warn > gpurun_out/r2_quickbench_collocation.log; cat gpurun_out/r2_quickbench_collocation.log
timeout 300 python tools/quickbench_shooting.py 2>&1 | grep -v Warn > gpurun_out/r2_quickbench_shooting.log
timeout 300 python tools/quickbench_node.py 2>&1 | grep -v Warn > gpurun_out/r2_quickbench_node_c5.log
MYR_LIB=$PWD/build/lib_ph.so timeout 300 python tools/phase_profile.py 1024 > gpurun_out/r2_phase_cycles.log 2>&1; cat gpurun_out/r2_phase_cycles.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -f -o gpurun_out/r2_ipm_trap_B1024 python tools/profile_run.py trap 1024 ipm > gpurun_out/prof1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -f -o gpurun_out/r2_ipm_hs_B148 python tools/profile_run.py hs 148 ipm > gpurun_out/prof2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -f -o gpurun_out/r2_ipm_node_B148 python tools/profile_node_ipm.py 148 > gpurun_out/prof3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for tool in racecheck memcheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.log; tail -2 gpurun_out/r2_sanitizer_$tool.log
done
ls -la gpurun_out/*.ncu-rep | tail -4; python tools/hs_fail_debug.py hs 8192 > gpurun_out/r2_hs_B8192_status.log 2>&1; tail -2 gpurun_out/r2_hs_B8192_status.log
