"""FBSM sweep kernel: solves/s and HBM roofline on one GPU (CUDA events on the launching stream, L2 flushed between timed
launches).  Algorithmic bytes per instance-sweep = (N + 1) * 8 * (4 n + 3 m) (csrc/fbsm.cuh); peak = MEASURED_PEAKS.json.

    python tools/quickbench_fbsm.py [profile SYSTEM N B]     # "profile": one warm-up + one launch, for ncu
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from myriad_b200.config import Config, HParams, OptimizerType
from myriad_b200.systems import SystemType
from myriad_b200.trajectory_optimizers import get_optimizer

def peak_gbs():
  try:
    d = json.load(open("MEASURED_PEAKS.json"))
    for k in ("hbm_gbs_sustained", "hbm_gbs", "hbm_copy_gbs", "hbm_gbps"):
      if k in d: return float(d[k]), k
    for k, v in d.items():
      if "hbm" in k.lower() and isinstance(v, (int, float)): return float(v), k
  except Exception:
    pass
  return 6536.7, "fallback (round-1 measured copy bandwidth)"

def run(name, N, B, reps=3):
  hp = HParams(system=SystemType[name], optimizer=OptimizerType.FBSM, fbsm_intervals=N)
  opt = get_optimizer(hp, Config(verbose=False, plot=False), hp.system())
  rng = np.random.Generator(np.random.PCG64(3))
  x0d = np.asarray(opt.system.x_0, dtype=np.float64)
  x0 = torch.as_tensor(x0d * (1 + 0.1 * rng.uniform(-1, 1, size=(B, x0d.shape[0])))).cuda()
  flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
  r = opt.solve_batch(x0); torch.cuda.synchronize()
  ms = []
  for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = opt.solve_batch(x0); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
  t = float(np.median(ms))
  it = r["iters"].cpu().numpy(); st = r["status"].cpu().numpy()
  n, m = x0d.shape[0], opt.system.control_size
  warp_sweeps = it.reshape(-1)[: B // 32 * 32].reshape(-1, 32).max(1).sum() * 32 + it[B // 32 * 32:].sum() if B >= 32 else it.sum()
  bytes_alg = float(it.sum()) * (N + 1) * 8 * (4 * n + 3 * m) + B * (N + 1) * 8 * (2 * n + m)  # + the initial fill
  pk, src = peak_gbs()
  print(f"FBSM {name} N={N} B={B}: {t:.2f} ms -> {B / t * 1e3:.0f} solves/s; sweeps {it.min()}/{np.median(it):.0f}/{it.max()}; "
        f"solved {(st == 0).sum()}/{B}; algorithmic {bytes_alg / 1e9:.2f} GB -> {bytes_alg / t / 1e6:.0f} GB/s = "
        f"{bytes_alg / t / 1e6 / pk:.3f} of {pk:.0f} ({src})", flush=True)

def cpu_lines(name, N, B):
  """same sweep templates on the host cores (OpenMP over instances) and the NumPy oracle on one core, bounded samples"""
  from oracle import fbsm as OF
  hp = HParams(system=SystemType[name], optimizer=OptimizerType.FBSM, fbsm_intervals=N)
  opt = get_optimizer(hp, Config(verbose=False, plot=False), hp.system())
  rng = np.random.Generator(np.random.PCG64(3))
  x0d = np.asarray(opt.system.x_0, dtype=np.float64)
  x0 = x0d * (1 + 0.1 * rng.uniform(-1, 1, size=(B, x0d.shape[0])))
  t = time.time(); r = opt.host_solve_batch(x0); dt = time.time() - t
  print(f"FBSM {name} N={N}: host build of the same templates, {os.cpu_count()} cores, {B} instances: {dt:.2f} s -> {B / dt:.0f} solves/s", flush=True)
  t = time.time(); k = 0
  while time.time() - t < 10 and k < 8:
    OF.solve(name, N, x0[k]); k += 1
  dt = time.time() - t
  print(f"FBSM {name} N={N}: NumPy oracle (the reference's algorithm as the reference runs it, one core), {k} instances: {dt / k:.2f} s per solve -> {k / dt:.2f} solves/s", flush=True)

if __name__ == "__main__":
  if len(sys.argv) > 1 and sys.argv[1] == "profile":
    run(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), reps=1)
  else:
    for B in (1024, 16384, 65536, 262144):
      run("CANCERTREATMENT", 1000, B)
    run("SIMPLECASE", 1000, 65536)
    run("HIVTREATMENT", 1000, 65536)
    run("BEARPOPULATIONS", 1000, 65536)
    run("PREDATORPREY", 1000, 16384)
    cpu_lines("CANCERTREATMENT", 1000, 2048)
