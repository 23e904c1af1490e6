#!/bin/bash
# supplementary scaling point: 8192 instances per GPU (the per-rank tail of rare 45+-iteration instances is amortised)
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --batch 8192 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b8192_${N}gpu.json 2> gpurun_out/r2_bench_b8192_${N}gpu.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --batch 8192 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b8192_${N}gpu.json 2> gpurun_out/r2_bench_b8192_${N}gpu.err
fi
python - <<PY
import json
for ln in open("gpurun_out/r2_bench_b8192_${N}gpu.json"):
  if ln.startswith("{"):
    d = json.loads(ln); print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "e2e", "solved", "instances")}); print([(r["rank"], round(r["solve_ms"], 2), round(r["gather_ms"], 2), r["max_iters"]) for r in d["per_rank"]])
PY
tail -2 gpurun_out/r2_bench_b8192_${N}gpu.err
