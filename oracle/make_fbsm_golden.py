"""Generate tests/golden/fbsm_*.npz by running the UNMODIFIED reference FBSM
(/root/reference/myriad/trajectory_optimizers/forward_backward_sweep.py) under oracle/refshim.  Build container only:

    python -m oracle.make_fbsm_golden

Per system: the reference's solution {'x', 'u', 'adj'} for a few start states (the constructor's x_0 and seeded
perturbations of it) at fbsm_intervals = N, plus the number of sweeps the reference performed (counted by wrapping
``stopping_criterion``, which it calls once per sweep: forward_backward_sweep.py:94).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import refshim

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# system -> (fbsm_intervals, relative perturbations applied to x_0 for the extra start states)
CASES = {
  "SIMPLECASE": (1000, 2), "SIMPLECASEWITHBOUNDS": (200, 2), "CANCERTREATMENT": (1000, 2), "MOULDFUNGICIDE": (200, 1),
  "BIOREACTOR": (200, 2), "GLUCOSE": (200, 2), "HARVEST": (200, 2), "TIMBERHARVEST": (200, 1), "EPIDEMICSEIRN": (200, 1),
  "HIVTREATMENT": (200, 1), "BACTERIA": (200, 1), "PREDATORPREY": (200, 0), "BEARPOPULATIONS": (200, 1),
  "INVASIVEPLANT": (10, 2),  # discrete: N = int(T) whatever fbsm_intervals says (forward_backward_sweep.py:33-35)
}


def start_states(x0: np.ndarray, extra: int, seed: int) -> np.ndarray:
  rng = np.random.Generator(np.random.PCG64(seed))
  rows = [x0]
  for _ in range(extra):
    rows.append(x0 * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, size=x0.shape)))
  return np.stack(rows)


def run_reference(name: str, N: int, x0: np.ndarray):
  from myriad.config import Config, HParams, OptimizerType
  from myriad.systems import SystemType
  from myriad.trajectory_optimizers import get_optimizer
  hp = HParams(system=SystemType[name], optimizer=OptimizerType.FBSM, fbsm_intervals=N)
  cfg = Config(verbose=False, plot=False)
  system = hp.system()
  system.x_0 = refshim._wrap(np.array(x0, dtype=np.float64))  # FBSM.__init__ builds x_guess from system.x_0
  opt = get_optimizer(hp, cfg, system)
  sweeps = [0]
  inner = opt.stopping_criterion

  def counted(*a, **k):
    sweeps[0] += 1
    return inner(*a, **k)
  opt.stopping_criterion = counted
  sol = opt.solve()
  return {k: np.asarray(v, dtype=np.float64) for k, v in sol.items()}, sweeps[0]


def main():
  refshim.install()
  os.makedirs(GOLD, exist_ok=True)
  only = sys.argv[1:] or list(CASES)
  for name in only:
    N, extra = CASES[name]
    from myriad.systems import SystemType
    x0_default = np.asarray(SystemType[name].value().x_0, dtype=np.float64)
    x0s = start_states(x0_default, extra, seed=2021)
    t = time.time()
    xs, us, adjs, sweeps = [], [], [], []
    for x0 in x0s:
      sol, k = run_reference(name, N, x0)
      xs.append(sol["x"]); us.append(sol["u"]); adjs.append(sol["adj"]); sweeps.append(k)
    np.savez_compressed(os.path.join(GOLD, f"fbsm_{name.lower()}.npz"), N=np.int64(N), x0=x0s, x=np.stack(xs), u=np.stack(us),
                        adj=np.stack(adjs), sweeps=np.asarray(sweeps, dtype=np.int64))
    print(f"fbsm {name}: N={N} starts={len(x0s)} sweeps={sweeps} ({time.time() - t:.1f}s)", flush=True)


if __name__ == "__main__":
  sys.exit(main())
