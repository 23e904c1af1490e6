#!/bin/bash
# tools/gpu_multi.sh N: bench.py + C4 sweep on N GPUs of one box (gpurun --gpus N)
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
  timeout 600 python tools/c4_sweep.py > gpurun_out/r2_c4_sweep_${N}gpu.log 2>&1
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/c4_sweep.py > gpurun_out/r2_c4_sweep_${N}gpu.log 2>&1
fi
python - <<PY
import json
for ln in open("gpurun_out/r2_bench_${N}gpu.json"):
  if ln.startswith("{"):
    d = json.loads(ln); print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "e2e", "solved", "instances")}); print(d["per_rank"])
PY
tail -3 gpurun_out/r2_bench_${N}gpu.err
grep "^C4\|^GPUs" gpurun_out/r2_c4_sweep_${N}gpu.log
