"""BASELINE config C4: CANCERTREATMENT shooting 1x100, max_iter=500, batch sweep 2^6 .. 2^18 (TOTAL instances) on
1 / 2 / 4 / 8 GPUs: instances sharded by rows over the ranks, one NCCL all_gather of the packed solutions per solve.

    python tools/c4_sweep.py                                          # 1 GPU
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/c4_sweep.py
Time = max over ranks (CUDA events); one line per batch size from rank 0."""
import os, sys, torch, numpy as np
sys.path.insert(0, '.')
import torch.distributed as dist
from myriad_b200 import problems as PR
from myriad_b200.distributed import gather_solutions, pack_solution, shard_range
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
  dist.init_process_group("nccl", device_id=dev)
tr = PR.Transcription(SystemType.CANCERTREATMENT(), PR.SHOOTING, "HEUN", 1, 100)
eng = Engine(tr.desc())
if rank == 0:
  print(f"GPUs {world}; workspace: {eng.sizes.ipm_workspace_doubles} doubles/slot x min(B, {eng.sizes.ipm_workspace_slots}) slots", flush=True)
for e in range(6, 19, 2):
  Btot = 1 << e
  lo, hi = shard_range(Btot, rank, world)
  x0 = PR.sample_x0(tr.system, Btot)[lo:hi].to(dev)
  z0, lb, ub = PR.build_batch(tr, x0)
  zero = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
  def step(out=None):
    out = eng.ipm_solve(z0, lb, ub, max_iter=500, out=out)
    return out, gather_solutions(pack_solution(out["z"], out["lam"], out["obj"], zero, out["status"], out["iters"]))
  out, _ = step()
  torch.cuda.synchronize()
  if world > 1: dist.barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); out, allp = step(out); e1.record(); torch.cuda.synchronize()
  t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
  ok = torch.tensor([float((out["status"] == 0).sum())], dtype=torch.float64, device=dev)
  cm = torch.tensor([float(out["con_inf"][out["status"] == 0].max())], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(ok); dist.all_reduce(cm, op=dist.ReduceOp.MAX)
  if rank == 0:
    ms = float(t[0])
    print(f"C4 GPUs={world} B=2^{e}={Btot}: {ms:.2f} ms -> {Btot/ms*1e3:.0f} solves/s; solved {int(ok[0])}/{Btot}; cinf max (solved) {float(cm[0]):.1e}; "
          f"gathered {tuple(allp.shape)}; mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB", flush=True)
if world > 1:
  dist.barrier(); dist.destroy_process_group()
