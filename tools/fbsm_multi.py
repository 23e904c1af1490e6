"""FBSM weak scaling under torchrun: every rank sweeps its own B start states (no data-path collective), then ONE
all_gather of the optimal controls.  Device-timed (CUDA events), max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/fbsm_multi.py [SYSTEM N B]
"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from myriad_b200.config import Config, HParams, OptimizerType
from myriad_b200.systems import SystemType
from myriad_b200.trajectory_optimizers import get_optimizer

name = sys.argv[1] if len(sys.argv) > 1 else "CANCERTREATMENT"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
  dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hp = HParams(system=SystemType[name], optimizer=OptimizerType.FBSM, fbsm_intervals=N)
opt = get_optimizer(hp, Config(verbose=False, plot=False), hp.system())
rng = np.random.Generator(np.random.PCG64(100 + rank))
x0d = np.asarray(opt.system.x_0, dtype=np.float64)
x0 = torch.as_tensor(x0d * (1 + 0.1 * rng.uniform(-1, 1, size=(B, x0d.shape[0])))).cuda()
m = opt.system.control_size
gathered = torch.empty(world, opt.u_rows, m, B, dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def step():
  r = opt.solve_batch(x0)
  u = r["u"].permute(1, 2, 0).contiguous()  # the kernel's own [rows][m][B] storage
  if world > 1:
    dist.all_gather_into_tensor(gathered.view(world, -1), u.view(-1))
  return r

for _ in range(3):
  r = step()
torch.cuda.synchronize()
ms_solve, ms_all = [], []
for _ in range(5):
  flush.zero_()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
  e0.record(); r = opt.solve_batch(x0); e1.record()
  u = r["u"].permute(1, 2, 0).contiguous()
  if world > 1:
    dist.all_gather_into_tensor(gathered.view(world, -1), u.view(-1))
  e2.record(); torch.cuda.synchronize()
  ms_solve.append(e0.elapsed_time(e1)); ms_all.append(e0.elapsed_time(e2))
t = torch.tensor([float(np.median(ms_all)), float(np.median(ms_solve))], device="cuda")
if world > 1:
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
solved = torch.tensor([int((r["status"] == 0).sum())], device="cuda")
if world > 1:
  dist.all_reduce(solved)
if rank == 0:
  print(f"FBSM {name} N={N} GPUs={world} B/GPU={B}: step {t[0]:.2f} ms (sweeps {t[1]:.2f} ms, max over ranks) -> "
        f"{world * B / float(t[0]) * 1e3:.0f} solves/s; solved {int(solved)}/{world * B}; gathered u {tuple(gathered.shape)}", flush=True)
if world > 1:
  dist.barrier(); dist.destroy_process_group()
