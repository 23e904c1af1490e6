"""TrajectoryOptimizer: mirror of myriad/trajectory_optimizers/base.py:27-93.

Keeps the reference's fields (objective, constraints, bounds, guess, unravel, parametrized_*) and methods
(solve, solve_with_params) and adds ``solve_batch`` -- the capability the reference lacks (SURVEY.md headline
facts): many start states per launch.  objective / constraints / their derivatives are evaluated by the K1
CUDA kernel (myr_eval); they accept a flat decision vector and return NumPy values like the jitted reference
callables do.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch

from myriad_b200 import problems as PR
from myriad_b200.config import Config, HParams
from myriad_b200.systems import SystemType


class TrajectoryOptimizer(object):
  require_adj: bool = False

  def __init__(self, hp: HParams, cfg: Config, system, transcription: PR.Transcription):
    self.hp, self.cfg, self.system = hp, cfg, system
    self.transcription = transcription
    self._engine = None
    if hp.system == SystemType.INVASIVEPLANT:  # base.py:66-67
      raise NotImplementedError("Discrete systems are not compatible with Trajectory trajectory_optimizers")
    x0 = torch.as_tensor(np.asarray(system.x_0, dtype=np.float64)).reshape(1, -1).cuda()
    z0, lb, ub = PR.build_batch(transcription, x0)
    self.guess = z0[0].cpu().numpy()
    self.bounds = np.stack([lb[0].cpu().numpy(), ub[0].cpu().numpy()], axis=1)
    self.x_guess, self.u_guess = transcription.unravel(self.guess)
    nx = transcription.nx_nodes * transcription.n
    self.x_bounds, self.u_bounds = self.bounds[:nx], self.bounds[nx:]
    self.unravel: Callable = transcription.unravel
    if cfg.verbose:
      print("hp opt type", hp.optimizer)
      print("hp quadrature rule", hp.quadrature_rule)
      print(f"guess.shape = {self.guess.shape}")
      print(f"bounds.shape = {self.bounds.shape}")

  # ---- K1-backed callables (reference: closures objective / constraints + jax.grad / jax.jacrev)
  @property
  def engine(self):
    if self._engine is None:
      from myriad_b200.nlp_solvers import _engine_for
      self._engine = _engine_for(self.transcription)
    return self._engine

  def _eval(self, variables):
    z = torch.as_tensor(np.ascontiguousarray(variables), dtype=torch.float64).reshape(1, -1).cuda()
    return self.engine.eval(z)

  def objective(self, variables) -> float:
    return float(self._eval(variables).f[0])

  def constraints(self, variables) -> np.ndarray:
    return self._eval(variables).c[0].cpu().numpy()

  def objective_grad(self, variables) -> np.ndarray:
    return self._eval(variables).grad[0].cpu().numpy()

  def constraints_jac(self, variables) -> np.ndarray:
    return PR.dense_jacobian(self.transcription, self._eval(variables).Jblk)[0].cpu().numpy()

  def parametrized_objective(self, params, variables):
    return self._with_params(params).objective(variables)

  def parametrized_constraints(self, params, variables):
    return self._with_params(params).constraints(variables)

  def _with_params(self, params) -> "TrajectoryOptimizer":
    """params: constructor arguments of the physical system (useful_scripts.py:35), or -- for a NodeSystem -- the haiku
    parameter mapping of its MLP (NodeSystem.parametrized_dynamics, node_system.py:36-38)."""
    from myriad_b200.systems import NodeSystem
    if isinstance(self.system, NodeSystem):
      system = NodeSystem(params, self.system.true_system)
    else:
      system = self.hp.system(**params)
    cfg = Config(**{**self.cfg.__dict__, "verbose": False})
    return type(self)(self.hp, cfg, system)

  # ---- solves
  def _opt_inputs(self):
    return {'objective': self.objective, 'guess': self.guess, 'constraints': self.constraints, 'bounds': self.bounds,
            'unravel': self.unravel, 'transcription': self.transcription}

  def solve(self) -> Dict[str, np.ndarray]:
    from myriad_b200.nlp_solvers import solve
    return solve(self.hp, self.cfg, self._opt_inputs())

  def solve_with_params(self, params, guess: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """base.py:81-93: plan with a system built from ``params`` (hp.system(**params), useful_scripts.py:35)."""
    other = self._with_params(params)
    # like the reference, the guess and the bounds stay those of THIS optimizer; only the dynamics / cost change
    inputs = {**self._opt_inputs(), 'objective': other.objective, 'constraints': other.constraints,
              'transcription': other.transcription}
    if guess is not None:
      inputs['guess'] = guess
    from myriad_b200.nlp_solvers import solve
    return solve(self.hp, self.cfg, inputs)

  def solve_batch(self, x0s) -> Dict[str, torch.Tensor]:
    """Solve one NLP per row of ``x0s`` ([B, n], CUDA or host) in a single launch.  Returns device tensors:
    x [B, nx_nodes, n], u [B, nu_nodes, m], xs_and_us, cost, lambda, status, iters."""
    from myriad_b200.nlp_solvers import solve_batch
    x0s = torch.as_tensor(x0s, dtype=torch.float64).cuda().contiguous()
    z0, lb, ub = PR.build_batch(self.transcription, x0s)
    out = solve_batch(self.hp, self.cfg, self.transcription, z0, lb, ub)
    x, u = self.transcription.unravel(out["z"])
    return {'x': x, 'u': u, 'xs_and_us': out["z"], 'cost': out["obj"], 'lambda': out["lam"], 'status': out["status"],
            'iters': out["iters"], 'kkt_error': out["kkt_err"], 'constraint_violation': out["con_inf"]}
