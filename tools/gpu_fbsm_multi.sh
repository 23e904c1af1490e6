#!/bin/bash
# FBSM weak scaling on N GPUs of one box ($1 = N); the N=1 line is printed first on GPU 0
N=${1:-2}
mkdir -p gpurun_out
LOG=gpurun_out/r2_fbsm_scaling_${N}gpu.log
timeout 300 python tools/fbsm_multi.py 2>&1 | grep "^FBSM" > $LOG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/fbsm_multi.py 2>&1 | grep "^FBSM\|Error\|error" >> $LOG
cat $LOG
