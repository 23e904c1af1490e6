"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under
oracle/refshim.  Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_golden            # evaluation fixtures (seconds)
    python -m oracle.make_golden --solve    # + reference solve() with SLSQP (minutes)

Fixtures per case: guess, bounds, a seeded test point z, objective(z), constraints(z),
grad objective(z), dense jacobian of constraints(z) (reference closures + complex-step), and with
--solve the reference's own ``solve`` (myriad/nlp_solvers/__init__.py:18-98, SLSQP branch :50-52)
result plus ``get_state_trajectory_and_cost`` / ``get_defect`` (what run_trajectory_opt returns,
myriad/useful_scripts.py:47-49,76).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

from . import refshim

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (system, optimizer, quadrature, integration_method, intervals, cpi)
CASES = {
  # BASELINE.json configs (C1..C4)
  "c1_simplecase_shooting_10x100_heun": ("SIMPLECASE", "SHOOTING", "TRAPEZOIDAL", "HEUN", 10, 100),
  "c2_cartpole_trap_100": ("CARTPOLE", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 100, 1),
  "c2_cartpole_hs_100": ("CARTPOLE", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 100, 1),
  "c3_vanderpol_shooting_1x50_heun": ("VANDERPOL", "SHOOTING", "TRAPEZOIDAL", "HEUN", 1, 50),
  "c4_cancer_shooting_1x100_heun": ("CANCERTREATMENT", "SHOOTING", "TRAPEZOIDAL", "HEUN", 1, 100),
  # the reference's own smoke matrix (tests/tests.py:46-211), shrunk
  "t_simplecase_shooting_1x50_heun": ("SIMPLECASE", "SHOOTING", "TRAPEZOIDAL", "HEUN", 1, 50),
  "t_simplecase_shooting_20x3_heun": ("SIMPLECASE", "SHOOTING", "TRAPEZOIDAL", "HEUN", 20, 3),
  "t_simplecase_shooting_50x1_euler": ("SIMPLECASE", "SHOOTING", "TRAPEZOIDAL", "EULER", 50, 1),
  "t_simplecase_shooting_50x1_midpoint": ("SIMPLECASE", "SHOOTING", "TRAPEZOIDAL", "MIDPOINT", 50, 1),
  "t_simplecase_shooting_50x1_rk4": ("SIMPLECASE", "SHOOTING", "TRAPEZOIDAL", "RK4", 50, 1),
  "t_simplecase_trap_50": ("SIMPLECASE", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 50, 1),
  "t_simplecase_hs_50": ("SIMPLECASE", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 50, 1),
  # other systems x transcriptions at small sizes
  "s_vanderpol_trap_20": ("VANDERPOL", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 20, 1),
  "s_vanderpol_hs_10": ("VANDERPOL", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 10, 1),
  "s_vanderpol_shooting_4x5_rk4": ("VANDERPOL", "SHOOTING", "TRAPEZOIDAL", "RK4", 4, 5),
  "s_cartpole_shooting_5x4_heun": ("CARTPOLE", "SHOOTING", "TRAPEZOIDAL", "HEUN", 5, 4),
  "s_cartpole_shooting_3x4_rk4": ("CARTPOLE", "SHOOTING", "TRAPEZOIDAL", "RK4", 3, 4),
  "s_cartpole_trap_10": ("CARTPOLE", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "s_cancer_trap_20": ("CANCERTREATMENT", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 20, 1),
  "s_cancer_hs_10": ("CANCERTREATMENT", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 10, 1),
  "s_cancer_shooting_2x10_midpoint": ("CANCERTREATMENT", "SHOOTING", "TRAPEZOIDAL", "MIDPOINT", 2, 10),
  "s_cancer_shooting_2x10_euler": ("CANCERTREATMENT", "SHOOTING", "TRAPEZOIDAL", "EULER", 2, 10),
  # further SystemType members with generated device code (incl. the two time-dependent running costs)
  "x_mould_trap_10": ("MOULDFUNGICIDE", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "x_bioreactor_hs_8": ("BIOREACTOR", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 8, 1),
  "x_scwb_shooting_4x5_heun": ("SIMPLECASEWITHBOUNDS", "SHOOTING", "TRAPEZOIDAL", "HEUN", 4, 5),
  "x_glucose_trap_10": ("GLUCOSE", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "x_harvest_trap_10": ("HARVEST", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "x_harvest_shooting_2x8_rk4": ("HARVEST", "SHOOTING", "TRAPEZOIDAL", "RK4", 2, 8),
  "x_timber_hs_6": ("TIMBERHARVEST", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 6, 1),
  "x_timber_shooting_3x6_midpoint": ("TIMBERHARVEST", "SHOOTING", "TRAPEZOIDAL", "MIDPOINT", 3, 6),
  "x_seir_trap_10": ("SEIR", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "x_seir_shooting_2x5_heun": ("SEIR", "SHOOTING", "TRAPEZOIDAL", "HEUN", 2, 5),
  "x_seirn_hs_6": ("EPIDEMICSEIRN", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 6, 1),
  "x_hiv_trap_10": ("HIVTREATMENT", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "x_hiv_shooting_4x25_rk4": ("HIVTREATMENT", "SHOOTING", "TRAPEZOIDAL", "RK4", 4, 25),
  # systems with a terminal cost (trapezoid and shooting add it, Hermite-Simpson ignores it: SURVEY 9-5)
  "x_bacteria_trap_10": ("BACTERIA", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "x_bacteria_shooting_3x5_heun": ("BACTERIA", "SHOOTING", "TRAPEZOIDAL", "HEUN", 3, 5),
  "x_bacteria_hs_6": ("BACTERIA", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 6, 1),
  "x_tumour_trap_10": ("TUMOUR", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "x_tumour_shooting_2x6_rk4": ("TUMOUR", "SHOOTING", "TRAPEZOIDAL", "RK4", 2, 6),
  "x_predprey_shooting_100x1_heun": ("PREDATORPREY", "SHOOTING", "TRAPEZOIDAL", "HEUN", 100, 1),
  # two controls (the control-bound rows are filled control-major like in the reference: SURVEY 9-2)
  "x_bear_trap_10": ("BEARPOPULATIONS", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "x_bear_hs_6": ("BEARPOPULATIONS", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 6, 1),
  "x_bear_shooting_3x4_heun": ("BEARPOPULATIONS", "SHOOTING", "TRAPEZOIDAL", "HEUN", 3, 4),
  "x_bear_shooting_2x3_rk4": ("BEARPOPULATIONS", "SHOOTING", "TRAPEZOIDAL", "RK4", 2, 3),
  # round 2: the remaining continuous SystemType members -- ROCKETLANDING (n = 6, m = 2) and the two non-smooth ones
  # (clip / angle_normalize / jax.grad inside the dynamics: SURVEY 9-14)
  "y_rocket_trap_10": ("ROCKETLANDING", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "y_rocket_hs_5": ("ROCKETLANDING", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 5, 1),
  "y_rocket_shooting_3x4_heun": ("ROCKETLANDING", "SHOOTING", "TRAPEZOIDAL", "HEUN", 3, 4),
  "y_pendulum_trap_20": ("PENDULUM", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 20, 1),
  "y_pendulum_shooting_50x1_heun": ("PENDULUM", "SHOOTING", "TRAPEZOIDAL", "HEUN", 50, 1),
  "y_mountaincar_trap_20": ("MOUNTAINCAR", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 20, 1),
  "y_mountaincar_shooting_4x5_rk4": ("MOUNTAINCAR", "SHOOTING", "TRAPEZOIDAL", "RK4", 4, 5),
  # BASELINE config C5: CARTPOLE with neural-ODE MLP dynamics (3 x 64), planned the way the reference's
  # plan_with_node_model does (myriad/utils.py:230-242: system.dynamics <- net.apply(params, append(x, u))).
  # Weights: tests/golden/node_cartpole_64x64x64.npz (tools/fit_node.py).
  "c5_node_cartpole_trap_100": ("NODE_CARTPOLE", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 100, 1),
  "n_node_cartpole_trap_10": ("NODE_CARTPOLE", "COLLOCATION", "TRAPEZOIDAL", "HEUN", 10, 1),
  "n_node_cartpole_hs_6": ("NODE_CARTPOLE", "COLLOCATION", "HERMITE_SIMPSON", "RK4", 6, 1),
  "n_node_cartpole_shooting_3x4_heun": ("NODE_CARTPOLE", "SHOOTING", "TRAPEZOIDAL", "HEUN", 3, 4),
}

NODE_WEIGHTS = {"NODE_CARTPOLE": "node_cartpole_64x64x64.npz"}


def load_node_weights(sysname):
  """-> [(w, b), ...] in haiku layer order (linear, linear_1, ...: create_node.py:124-131)"""
  d = dict(np.load(os.path.join(GOLD, NODE_WEIGHTS[sysname])))
  out, i = [], 0
  while ("linear" if i == 0 else f"linear_{i}") + "/w" in d:
    k = "linear" if i == 0 else f"linear_{i}"
    out.append((d[k + "/w"], d[k + "/b"]))
    i += 1
  return out


def _haiku_mlp_apply(weights, x_and_u):
  """net.apply(params, x_and_u) of create_node.py:110-117 restated (haiku is not installed): hk.Linear is x @ w + b,
  jax.nn.sigmoid between layers, linear output."""
  h = x_and_u
  for i, (w, b) in enumerate(weights):
    h = h @ w + b
    if i + 1 < len(weights):
      h = 1.0 / (1.0 + np.exp(-h))
  return h

# which cases also get a reference SLSQP solve (kept to what finishes in minutes)
SOLVE_CASES = [
  "c2_cartpole_trap_100", "c3_vanderpol_shooting_1x50_heun", "c4_cancer_shooting_1x100_heun",
  "c1_simplecase_shooting_10x100_heun", "t_simplecase_shooting_1x50_heun", "t_simplecase_trap_50",
  "t_simplecase_hs_50", "s_vanderpol_trap_20", "s_cancer_trap_20", "s_cartpole_trap_10",
  "t_simplecase_shooting_20x3_heun", "s_vanderpol_hs_10", "n_node_cartpole_trap_10",
  "x_mould_trap_10", "x_glucose_trap_10", "x_seir_trap_10", "x_hiv_trap_10", "x_bacteria_trap_10", "x_bacteria_shooting_3x5_heun", "x_tumour_trap_10", "x_predprey_shooting_100x1_heun", "x_bear_trap_10", "x_bear_shooting_3x4_heun",  "x_harvest_trap_10", "x_scwb_shooting_4x5_heun",
  # full-size BASELINE configs whose reference solve takes tens of minutes under the shim (generated once, round 2)
  "c2_cartpole_hs_100", "c5_node_cartpole_trap_100",
  # (ROCKETLANDING and PENDULUM: the reference's own SLSQP ends infeasible -- |c| = 187 / 0.07 -- on these: with m = 2 the
  #  trapezoid pins node N-1 instead of N and scrambles the control bounds; PENDULUM's target angle pi
  #  sits on the discontinuity of angle_normalize.  They are covered by the evaluation / rollout fixtures only.)
  "y_mountaincar_trap_20", "y_mountaincar_shooting_4x5_rk4",
]


def _ref_objects(case):
  from myriad.config import (Config, HParams, IntegrationMethod, NLPSolverType, OptimizerType, QuadratureRule)
  from myriad.systems import SystemType
  from myriad.trajectory_optimizers import get_optimizer
  sysname, opt, quad, meth, intervals, cpi = CASES[case]
  node = sysname.startswith("NODE_")
  true_name = sysname[5:] if node else sysname
  hp = HParams(system=SystemType[true_name], optimizer=OptimizerType[opt], nlpsolver=NLPSolverType.SLSQP,
               integration_method=IntegrationMethod[meth], quadrature_rule=QuadratureRule[quad],
               intervals=intervals, controls_per_interval=cpi, max_iter=1000)
  cfg = Config(verbose=False, plot=False)
  system = hp.system()
  import io, contextlib
  with contextlib.redirect_stdout(io.StringIO()):  # shooting.py:53 prints
    optimizer = get_optimizer(hp, cfg, system)
  if node:  # exactly what plan_with_node_model does before calling optimizer.solve() (myriad/utils.py:231-236)
    import jax.numpy as jnp
    weights = load_node_weights(sysname)
    system.dynamics = lambda x, u, t=None: _haiku_mlp_apply(weights, jnp.append(x, u))
  return hp, cfg, system, optimizer


def test_point(guess: np.ndarray, bounds: np.ndarray, seed: int) -> np.ndarray:
  """A seeded interior point near the guess (the guess itself has u == 0, which hides terms)."""
  rng = np.random.Generator(np.random.PCG64(seed))
  z = guess + 0.05 * rng.standard_normal(guess.shape) * np.maximum(1.0, np.abs(guess))
  lo, hi = bounds[:, 0], bounds[:, 1]
  fin = np.isfinite(lo) & np.isfinite(hi)
  width = np.where(fin, hi - lo, 1.0)
  z = np.where(fin, np.clip(z, lo + 0.05 * width, hi - 0.05 * width), z)
  z = np.where(lo == hi, lo, z)
  return z


def make_eval_fixture(case: str) -> dict:
  import jax
  hp, cfg, system, optimizer = _ref_objects(case)
  guess = np.asarray(optimizer.guess, dtype=np.float64)
  bounds = np.asarray(optimizer.bounds, dtype=np.float64)
  z = test_point(guess, bounds, seed=2019)
  out = {
    "guess": guess, "bounds": bounds, "z": z,
    "obj_guess": np.float64(optimizer.objective(guess)),
    "con_guess": np.asarray(optimizer.constraints(guess), dtype=np.float64),
    "obj_z": np.float64(optimizer.objective(z)),
    "con_z": np.asarray(optimizer.constraints(z), dtype=np.float64),
    "grad_z": np.asarray(jax.grad(optimizer.objective)(z), dtype=np.float64),
    "jac_z": np.asarray(jax.jacrev(optimizer.constraints)(z), dtype=np.float64),
  }
  # post-solve rollout of the TRUE system under the test-point controls (myriad/utils.py:258-324)
  from myriad.utils import get_defect, get_state_trajectory_and_cost
  _, u = optimizer.unravel(z)
  if CASES[case][0].startswith("NODE_"):
    system = hp.system()  # the verification rollout integrates the TRUE system (useful_scripts.py:43-49)
  try:
    xs, c = get_state_trajectory_and_cost(hp, system, system.x_0, u)
    out["rollout_states"] = np.asarray(xs, dtype=np.float64)
    out["rollout_cost"] = np.float64(np.squeeze(c))
    d = get_defect(system, xs)
    if d is not None:
      out["rollout_defect"] = np.asarray(d, dtype=np.float64)
  except IndexError:
    # numpy raises where jax clamps: HS + non-RK4 etc.
    pass
  return out


def make_solve_fixture(case: str) -> dict:
  from myriad.nlp_solvers import solve
  from myriad.utils import get_defect, get_state_trajectory_and_cost
  hp, cfg, system, optimizer = _ref_objects(case)
  t = time.time()
  opt_inputs = {"objective": optimizer.objective, "guess": optimizer.guess, "constraints": optimizer.constraints,
                "bounds": optimizer.bounds, "unravel": optimizer.unravel}
  res = solve(hp, cfg, opt_inputs)  # reference code, SLSQP branch
  secs = time.time() - t
  out = {"sol_x": np.asarray(res["x"]), "sol_u": np.asarray(res["u"]), "sol_z": np.asarray(res["xs_and_us"]),
         "sol_cost": np.float64(res["cost"]), "sol_seconds": np.float64(secs),
         "sol_con_inf": np.float64(np.abs(optimizer.constraints(res["xs_and_us"])).max())}
  if CASES[case][0].startswith("NODE_"):
    system = hp.system()
  try:
    xs, c = get_state_trajectory_and_cost(hp, system, system.x_0, res["u"])
    out["sol_rollout_cost"] = np.float64(np.squeeze(c))
    d = get_defect(system, xs)
    if d is not None:
      out["sol_rollout_defect"] = np.asarray(d, dtype=np.float64)
  except IndexError:
    pass
  return out


def make_integrator_kat() -> dict:
  """tests/tests.py:19-43: RK4 on y' = y, y(0)=1, 99 steps over linspace(0,1,100) (+ the other
  three methods through the same reference function)."""
  from myriad.config import IntegrationMethod
  from myriad.utils import integrate
  N = 100
  t = np.linspace(0., 1., N)
  h = t[1]
  out = {}
  for meth in IntegrationMethod:
    us = np.concatenate([t, np.full(N + 1, t[-1])])  # padded so numpy indexing == jax clamping
    _, states = integrate(lambda s, c, tt: s, np.array([1.]), us, h, N - 1, t, integration_method=meth)
    out[meth.name] = np.asarray(states, dtype=np.float64)
  return out


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--solve", action="store_true")
  ap.add_argument("--only", default=None)
  args = ap.parse_args()
  refshim.install()
  os.makedirs(GOLD, exist_ok=True)
  if not args.only:
    np.savez(os.path.join(GOLD, "integrator_kat.npz"), **make_integrator_kat())
  for case in CASES:
    if args.only and args.only != case:
      continue
    path = os.path.join(GOLD, case + ".npz")
    t = time.time()
    fx = make_eval_fixture(case)
    if args.solve and case in SOLVE_CASES:
      fx.update(make_solve_fixture(case))
    elif os.path.exists(path):  # keep an earlier solve
      old = dict(np.load(path))
      fx.update({k: v for k, v in old.items() if k.startswith("sol_")})
    np.savez_compressed(path, **fx)
    print(f"{case}: nvars={fx['guess'].shape[0]} ncon={fx['con_z'].shape[0]} "
          f"{'solved cost=%.10f ' % fx['sol_cost'] if 'sol_cost' in fx else ''}({time.time() - t:.1f}s)", flush=True)


if __name__ == "__main__":
  sys.exit(main())
