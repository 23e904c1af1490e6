"""CPU oracle: derivatives and NLP solve, restating myriad/nlp_solvers/__init__.py:18-98.

TEST INFRASTRUCTURE ONLY (see oracle/systems.py header).

The reference wraps objective/constraints in jax.grad / jax.jacrev (:31-42) and hands them to
cyipopt / SciPy.  IPOPT is not available in this image, so the oracle uses the SciPy solvers the
reference itself supports (SLSQP :50-52 -- the solver all of the reference's own tests use,
tests/tests.py:53 -- and trust-constr :53-55).  First derivatives are exact complex-step
derivatives evaluated in one batched call; second derivatives (only needed to check the CUDA
Hessian blocks) come from torch.func.
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np
from scipy.optimize import minimize

from .transcription import Transcription, get_defect, get_state_trajectory_and_cost

_CS = 1e-30


def _perturbed(z):
  nv = z.shape[0]
  zc = np.broadcast_to(z.astype(np.complex128), (nv, nv)).copy()
  zc[np.arange(nv), np.arange(nv)] += 1j * _CS
  return zc


def objective_grad(tr: Transcription, z: np.ndarray) -> np.ndarray:
  """jax.grad(objective) (nlp_solvers/__init__.py:40)"""
  return np.imag(tr.objective(_perturbed(z))) / _CS


def constraints_jac(tr: Transcription, z: np.ndarray) -> np.ndarray:
  """jax.jacrev(constraints) (nlp_solvers/__init__.py:37): dense (ncon, nvars)"""
  return (np.imag(tr.constraints(_perturbed(z))) / _CS).T.copy()


def lagrangian_hessian(tr: Transcription, z: np.ndarray, lam: np.ndarray, obj_factor: float = 1.0) -> np.ndarray:
  """Dense Hessian of  obj_factor * f(z) + lam . c(z)  via torch.func (fp64)."""
  import torch

  def lag(zz):
    return obj_factor * tr.objective(zz) + (torch.as_tensor(lam) * tr.constraints(zz)).sum()

  H = torch.func.hessian(lag)(torch.as_tensor(z, dtype=torch.float64))
  return H.numpy()


def solve(tr: Transcription, nlpsolver: str = "SLSQP", max_iter: int = 1000, guess=None,
          ftol: float = 1e-6, verbose: bool = False) -> Dict[str, np.ndarray]:
  """myriad/nlp_solvers/__init__.py:18-98.  Result dict keys as :90-96.

  ``ftol`` is SciPy's SLSQP default (1e-6) unless tightened; the reference passes only
  ``maxiter`` (:41)."""
  x0 = tr.guess if guess is None else guess
  fun = lambda z: float(tr.objective(z))
  jac = lambda z: objective_grad(tr, z)
  cons = {"type": "eq", "fun": lambda z: tr.constraints(z), "jac": lambda z: constraints_jac(tr, z)}
  t1 = time.time()
  if nlpsolver == "SLSQP":
    sol = minimize(fun, x0, method="SLSQP", jac=jac, constraints=cons, bounds=tr.bounds,
                   options={"maxiter": max_iter, "ftol": ftol})
    lam = None
  elif nlpsolver == "TRUST":
    sol = minimize(fun, x0, method="trust-constr", jac=jac, constraints=cons, bounds=tr.bounds,
                   options={"maxiter": max_iter})
    lam = sol["v"]
  else:
    raise ValueError(nlpsolver)
  t2 = time.time()
  x, u = tr.unravel(sol["x"])
  res = {"x": x, "u": u, "xs_and_us": sol["x"], "cost": sol["fun"], "success": bool(sol["success"]),
         "nit": int(sol.get("nit", -1)), "seconds": t2 - t1}
  if lam is not None:
    res["lambda"] = lam
  if verbose:
    print("Solver exited with success:", sol["success"], f"in {t2 - t1:.2f}s, nit={res['nit']}")
  return res


def run_trajectory_opt(system, tr: Transcription, intervals: int, cpi: int, method: str, **solve_kw):
  """myriad/useful_scripts.py:26-76: solve, then re-integrate the TRUE system under the solved
  controls; returns (cost, defect) like the reference, plus the solution dict."""
  sol = solve(tr, **solve_kw)
  opt_x, c = get_state_trajectory_and_cost(system, intervals, cpi, method, system.x_0, sol["u"])
  return float(c), get_defect(system, opt_x), sol
