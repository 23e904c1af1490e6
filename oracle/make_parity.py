"""Round-2 parity fixtures (TEST INFRASTRUCTURE; run in the build container only, /root/reference needed for part c):

    python -m oracle.make_parity tight      # (a) tight-tolerance solves -> tests/golden/tight_<case>.npz
    python -m oracle.make_parity random     # (b) 16 random-x0 rows of the bench workload -> tests/golden/random_x0_c2.npz
    python -m oracle.make_parity params     # (c) solve_with_params with non-default CARTPOLE parameters (reference code
                                            #     under oracle/refshim) -> tests/golden/params_*.npz

(a) The reference's own solve() runs SciPy's SLSQP with its default ftol=1e-6 (nlp_solvers/__init__.py:50-52), which
stops up to ~3e-5 short of the optimum on flat objectives, so trajectories of the committed sol_* fixtures agree with a
converged solver only to that flatness.  The tight fixtures re-solve the SAME NLP (oracle/transcription.py, pinned to
the reference's objective / constraints / derivatives by tests/test_oracle_vs_reference.py) with ftol=1e-12, so that
state / control trajectories, the re-integrated cost and the terminal defect (useful_scripts.py:47-49,76) can be
asserted against the CUDA path.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TIGHT_CASES = [
  "c2_cartpole_trap_100", "c3_vanderpol_shooting_1x50_heun", "c4_cancer_shooting_1x100_heun",
  "t_simplecase_shooting_1x50_heun", "t_simplecase_trap_50", "t_simplecase_hs_50", "t_simplecase_shooting_20x3_heun",
  "s_vanderpol_trap_20", "s_vanderpol_hs_10", "s_cancer_trap_20", "s_cartpole_trap_10", "n_node_cartpole_trap_10",
  "x_mould_trap_10", "x_seir_trap_10", "x_bear_trap_10", "x_harvest_trap_10", "x_bacteria_trap_10",
]

# non-default physical parameters for part (c): hp.system(**params) (useful_scripts.py:35)
CARTPOLE_PARAMS = {"g": 9.0, "m1": 1.2, "m2": 0.4, "length": 0.6}


def _oracle_tr(case):
  from .make_golden import CASES
  from .systems import make_system
  from .transcription import make_transcription
  sysname, opt, quad, meth, intervals, cpi = CASES[case]
  system = make_system(sysname)
  return system, make_transcription(system, opt, intervals, cpi, meth, quad), (sysname, opt, quad, meth, intervals, cpi)


def make_tight(case):
  from . import nlp
  from .systems import make_system
  from .transcription import get_defect, get_state_trajectory_and_cost
  system, tr, (sysname, opt, quad, meth, intervals, cpi) = _oracle_tr(case)
  t = time.time()
  r = nlp.solve(tr, "SLSQP", ftol=1e-12, max_iter=3000)
  out = {"x": r["x"], "u": r["u"], "z": r["xs_and_us"], "cost": np.float64(r["cost"]), "success": np.bool_(r["success"]),
         "nit": np.int64(r["nit"]), "con_inf": np.float64(np.abs(tr.constraints(r["xs_and_us"])).max()),
         "seconds": np.float64(time.time() - t)}
  true = make_system(sysname[5:]) if sysname.startswith("NODE_") else system
  try:
    xs, c = get_state_trajectory_and_cost(true, intervals, cpi, meth, true.x_0, r["u"])
    out["rollout_cost"] = np.float64(c)
    d = get_defect(true, xs)
    if d is not None:
      out["rollout_defect"] = np.asarray(d, dtype=np.float64)
  except IndexError:
    pass
  return out


def make_random(rows=16):
  """Rows of the bench workload (CARTPOLE trapezoid N=100, x0 = sample_x0(seed 2019)) solved by SLSQP at the reference's
  default ftol and tight."""
  from . import nlp
  from .cpu_baseline import sample_x0
  from .systems import make_system
  from .transcription import make_transcription
  import multiprocessing as mp
  system = make_system("CARTPOLE")
  x0s = sample_x0(system, rows)
  with mp.get_context("spawn").Pool(min(rows, os.cpu_count() or 1)) as pool:
    res = pool.map(_solve_row, [(x0s[i],) for i in range(rows)])
  return {"x0": x0s, "cost": np.array([r[0] for r in res]), "cost_tight": np.array([r[1] for r in res]),
          "u_tight": np.stack([r[2] for r in res]), "x_tight": np.stack([r[3] for r in res]),
          "success": np.array([r[4] for r in res])}


def _solve_row(args):
  from . import nlp
  from .systems import make_system
  from .transcription import make_transcription
  (x0,) = args
  system = make_system("CARTPOLE")
  system.x_0 = np.asarray(x0, dtype=np.float64)
  tr = make_transcription(system, "COLLOCATION", 100, 1, "HEUN", "TRAPEZOIDAL")
  a = nlp.solve(tr, "SLSQP")
  b = nlp.solve(tr, "SLSQP", ftol=1e-12, max_iter=3000)
  return float(a["cost"]), float(b["cost"]), b["u"], b["x"], bool(a["success"] and b["success"])


def make_params():
  """Reference code under the shim.  SHOOTING: the reference's own solve_with_params (base.py:81-93; the only
  transcription whose parametrized path works).  Trapezoid: the evident intent -- the optimizer
  of hp.system(**params) -- because trapezoidal.py:204-206 passes 7 in_axes for 5 arguments."""
  from . import refshim
  refshim.install()
  import contextlib
  import io
  from myriad.config import Config, HParams, IntegrationMethod, NLPSolverType, OptimizerType, QuadratureRule
  from myriad.systems import SystemType
  from myriad.trajectory_optimizers import get_optimizer
  from myriad.utils import get_state_trajectory_and_cost
  out = {}
  cfg = Config(verbose=False, plot=False)
  # (1) shooting 5 x 4 HEUN through solve_with_params
  hp = HParams(system=SystemType.CARTPOLE, optimizer=OptimizerType.SHOOTING, nlpsolver=NLPSolverType.SLSQP,
               integration_method=IntegrationMethod.HEUN, intervals=5, controls_per_interval=4, max_iter=1000)
  with contextlib.redirect_stdout(io.StringIO()):
    opt = get_optimizer(hp, cfg, hp.system())
  rng = np.random.Generator(np.random.PCG64(7))
  z = np.asarray(opt.guess) + 0.05 * rng.standard_normal(np.asarray(opt.guess).shape)
  out["shoot_z"] = z
  out["shoot_obj_z"] = np.float64(opt.parametrized_objective(CARTPOLE_PARAMS, z))
  out["shoot_con_z"] = np.asarray(opt.parametrized_constraints(CARTPOLE_PARAMS, z), dtype=np.float64)
  res = opt.solve_with_params(CARTPOLE_PARAMS)
  out["shoot_sol_z"] = np.asarray(res["xs_and_us"]); out["shoot_sol_cost"] = np.float64(res["cost"])
  out["shoot_sol_con_inf"] = np.float64(np.abs(opt.parametrized_constraints(CARTPOLE_PARAMS, res["xs_and_us"])).max())
  # (2) trapezoid N = 10 on hp.system(**params)
  hp2 = HParams(system=SystemType.CARTPOLE, optimizer=OptimizerType.COLLOCATION, nlpsolver=NLPSolverType.SLSQP,
                quadrature_rule=QuadratureRule.TRAPEZOIDAL, intervals=10, max_iter=1000)
  psys = hp2.system(**CARTPOLE_PARAMS)
  opt2 = get_optimizer(hp2, cfg, psys)
  z2 = np.asarray(opt2.guess) + 0.05 * rng.standard_normal(np.asarray(opt2.guess).shape)
  out["trap_z"] = z2
  out["trap_obj_z"] = np.float64(opt2.objective(z2))
  out["trap_con_z"] = np.asarray(opt2.constraints(z2), dtype=np.float64)
  res2 = opt2.solve()
  out["trap_sol_z"] = np.asarray(res2["xs_and_us"]); out["trap_sol_cost"] = np.float64(res2["cost"])
  out["trap_sol_con_inf"] = np.float64(np.abs(opt2.constraints(res2["xs_and_us"])).max())
  _, c = get_state_trajectory_and_cost(hp2, psys, psys.x_0, res2["u"])
  out["trap_sol_rollout_cost"] = np.float64(np.squeeze(c))
  out["params"] = np.array([CARTPOLE_PARAMS[k] for k in ("g", "m1", "m2", "length")])
  return out


def make_dataset():
  """Batched dataset rollouts (myriad/utils.py:422-424: integrate_time_independent_in_parallel over all trajectories of
  a dataset) by the reference code under the shim, for seeded random-walk controls, plus the reference's extragradient
  iterates (nlp_solvers/extra_gradient.py) on its own test problem (tests/tests.py:253-265)."""
  from . import refshim
  refshim.install()
  from myriad.config import IntegrationMethod
  from myriad.systems import SystemType
  from myriad.utils import integrate_time_independent_in_parallel
  out = {}
  rng = np.random.Generator(np.random.PCG64(11))
  for name, steps in (("CARTPOLE", 20), ("VANDERPOL", 30), ("ROCKETLANDING", 12)):
    system = SystemType[name]()
    n = system.x_0.shape[0]
    b = np.asarray(system.bounds, dtype=np.float64)
    m = b.shape[0] - n
    total = 6
    us = rng.uniform(b[n:, 0], b[n:, 1], (total, 2 * steps + 1, m)) * 0.5
    x0 = np.clip(np.asarray(system.x_0)[None] + 0.1 * rng.standard_normal((total, n)), b[:n, 0], b[:n, 1])
    out[f"{name}_us"] = us; out[f"{name}_x0"] = x0; out[f"{name}_steps"] = np.int64(steps)
    for meth in IntegrationMethod:
      u_in = us if meth == IntegrationMethod.RK4 else us[:, :steps + 1]
      u_call = u_in if m > 1 else u_in[..., 0]  # the reference squeezes scalar controls (shooting.py:128-129)
      _, xs = integrate_time_independent_in_parallel(system.dynamics, x0, u_call, system.T / steps, steps, meth)
      out[f"{name}_{meth.name}_xs"] = np.asarray(xs, dtype=np.float64)
  return out


def make_exgd():
  """The reference's extragradient solver (nlp_solvers/extra_gradient.py:10-82) run by the reference's own solve() on
  the problem of its own test (tests/tests.py:253-265: SIMPLECASE, SHOOTING 50 x 1, HEUN), 200 steps."""
  from . import refshim
  refshim.install()
  import contextlib
  import io
  from myriad.config import Config, HParams, IntegrationMethod, NLPSolverType, OptimizerType
  from myriad.nlp_solvers import solve
  from myriad.systems import SystemType
  from myriad.trajectory_optimizers import get_optimizer
  hp = HParams(system=SystemType.SIMPLECASE, optimizer=OptimizerType.SHOOTING, nlpsolver=NLPSolverType.EXTRAGRADIENT,
               integration_method=IntegrationMethod.HEUN, intervals=50, controls_per_interval=1, max_iter=20)
  assert hp.max_iter == 200  # config.py:100-101
  cfg = Config(verbose=False, plot=False)
  with contextlib.redirect_stdout(io.StringIO()):
    opt = get_optimizer(hp, cfg, hp.system())
    res = solve(hp, cfg, {"objective": opt.objective, "guess": opt.guess, "constraints": opt.constraints, "bounds": opt.bounds,
                          "unravel": opt.unravel})
  return {"z": np.asarray(res["xs_and_us"]), "lam": np.asarray(res["lambda"]), "cost": np.float64(res["cost"]),
          "guess": np.asarray(opt.guess), "bounds": np.asarray(opt.bounds)}


def main():
  what = sys.argv[1] if len(sys.argv) > 1 else "tight"
  only = sys.argv[2] if len(sys.argv) > 2 else None
  if what == "tight":
    for case in TIGHT_CASES:
      if only and only != case:
        continue
      fx = make_tight(case)
      np.savez_compressed(os.path.join(GOLD, "tight_" + case + ".npz"), **fx)
      print(f"{case}: cost {float(fx['cost']):.12f} success {bool(fx['success'])} nit {int(fx['nit'])} |c| {float(fx['con_inf']):.2e} ({float(fx['seconds']):.1f}s)", flush=True)
  elif what == "random":
    fx = make_random()
    np.savez_compressed(os.path.join(GOLD, "random_x0_c2.npz"), **fx)
    print("random rows:", fx["cost"], fx["cost_tight"], fx["success"])
  elif what == "dataset":
    fx = make_dataset()
    np.savez_compressed(os.path.join(GOLD, "dataset_rollouts.npz"), **fx)
    print({k: np.shape(v) for k, v in fx.items()})
  elif what == "exgd":
    fx = make_exgd()
    np.savez_compressed(os.path.join(GOLD, "exgd_simplecase.npz"), **fx)
    print("exgd cost", float(fx["cost"]), "lam[:3]", fx["lam"][:3])
  elif what == "params":
    fx = make_params()
    np.savez_compressed(os.path.join(GOLD, "params_cartpole.npz"), **fx)
    print({k: (v if np.ndim(v) == 0 else v.shape) for k, v in fx.items()})


if __name__ == "__main__":
  sys.exit(main())
