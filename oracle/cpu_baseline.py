"""CPU baseline runner: the reference's algorithm (oracle restatement, SciPy SLSQP -- the solver the
reference supports at myriad/nlp_solvers/__init__.py:50-52 and uses in all of its own tests) timed on host
cores over a bounded sample of the synthetic workload.  TEST/BENCH INFRASTRUCTURE (see oracle/systems.py).

    python -m oracle.cpu_baseline --system CARTPOLE --optimizer COLLOCATION --quadrature TRAPEZOIDAL \
        --intervals 100 --instances 8 --procs 8

prints one JSON line {"solves_per_s", "instances", "procs", "seconds", "costs", "success"}.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np


def sample_x0(system, B: int, seed: int = 2019, spread: float = 0.1) -> np.ndarray:
  """Same recipe as the GPU workload: x_0 + spread * N(0, I) clipped to the state bounds,
  torch CPU generator, row 0 unperturbed."""
  import torch
  g = torch.Generator(device="cpu").manual_seed(seed)
  noise = torch.randn(B, system.n, generator=g, dtype=torch.float64).numpy().copy()
  noise[0] = 0.0
  x0 = system.x_0[None, :] + spread * noise
  return np.clip(x0, system.bounds[:system.n, 0], system.bounds[:system.n, 1])


def _solve_one(args):
  os.environ["OMP_NUM_THREADS"] = "1"
  sysname, optimizer, quadrature, method, intervals, cpi, x0, max_iter = args
  try:
    import torch
    torch.set_num_threads(1)
  except Exception:
    pass
  from oracle import nlp
  from oracle.systems import make_system
  from oracle.transcription import make_transcription
  system = make_system(sysname)
  system.x_0 = np.asarray(x0, dtype=np.float64)
  tr = make_transcription(system, optimizer, intervals, cpi, method, quadrature)
  t = time.time()
  r = nlp.solve(tr, "SLSQP", max_iter=max_iter)
  return float(r["cost"]), bool(r["success"]), time.time() - t, int(r["nit"])


def run(sysname, optimizer, quadrature, method, intervals, cpi, instances, procs, max_iter=1000, seed=2019):
  from oracle.systems import make_system
  system = make_system(sysname)
  x0s = sample_x0(system, max(instances, 1), seed=seed)
  jobs = [(sysname, optimizer, quadrature, method, intervals, cpi, x0s[i], max_iter) for i in range(instances)]
  t = time.time()
  if procs > 1:
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
      res = pool.map(_solve_one, jobs)
  else:
    res = [_solve_one(j) for j in jobs]
  secs = time.time() - t
  return {"solves_per_s": instances / secs, "instances": instances, "procs": procs, "seconds": secs,
          "costs": [r[0] for r in res], "success": [r[1] for r in res], "per_solve_s": [r[2] for r in res],
          "nit": [r[3] for r in res]}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--system", default="CARTPOLE")
  ap.add_argument("--optimizer", default="COLLOCATION")
  ap.add_argument("--quadrature", default="TRAPEZOIDAL")
  ap.add_argument("--method", default="HEUN")
  ap.add_argument("--intervals", type=int, default=100)
  ap.add_argument("--cpi", type=int, default=1)
  ap.add_argument("--instances", type=int, default=0)
  ap.add_argument("--procs", type=int, default=0)
  ap.add_argument("--max_iter", type=int, default=1000)
  a = ap.parse_args()
  procs = a.procs or (os.cpu_count() or 1)
  inst = a.instances or procs
  print(json.dumps(run(a.system, a.optimizer, a.quadrature, a.method, a.intervals, a.cpi, inst, procs, a.max_iter)))


if __name__ == "__main__":
  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  main()
