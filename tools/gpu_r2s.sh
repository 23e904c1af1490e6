#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -c 1 -f -o gpurun_out/r2b_ipm_trap_B296 python tools/profile_run.py trap 296 ipm > gpurun_out/prof_r2b.log 2>&1
tail -3 gpurun_out/prof_r2b.log
