#!/bin/bash
# A/B of the sweep kernel's register cap (resident CTAs per SM the allocation must allow): default build vs MYR_FBSM_MIN_CTAS variants
mkdir -p gpurun_out
LOG=gpurun_out/r2_exp_fbsm_min_ctas.log; : > $LOG
for lib in myriad_b200/libmyriad_b200.so build/lib_fbsm_min20.so build/lib_fbsm_min24.so; do
  echo "== $lib" >> $LOG
  for args in "CANCERTREATMENT 1000 262144" "SIMPLECASE 1000 65536" "HIVTREATMENT 1000 65536"; do
    MYR_LIB=$PWD/$lib timeout 200 python tools/quickbench_fbsm.py profile $args 2>&1 | grep "^FBSM" | cut -c1-110 >> $LOG
  done
done
cat $LOG
