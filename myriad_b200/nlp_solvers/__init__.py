"""NLP solve: mirror of myriad/nlp_solvers/__init__.py:18-98 on top of the batched CUDA interior-point
solver (myr_ipm_solve), which replaces the cyipopt call at :56-58.

``solve(hp, cfg, opt_dict)`` keeps the reference's signature and result dictionary
(``x, u, xs_and_us, cost, lambda``, :90-96).  Objective / constraint *callables* cannot be traced into CUDA the way
jax traces them: the kernels are selected by a ``Transcription`` descriptor.  It is taken from ``opt_dict['transcription']``
when the optimizer put it there, from the optimizer the callables are bound to, or derived from ``hp`` for an ``opt_dict``
built the reference way (base.py:68-78) -- in that last case the user's callables are CHECKED against the kernels at the
guess, and a mismatch (custom or parametrized callables that are not this transcription's) raises instead of silently
solving a different problem.
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np
import torch

from myriad_b200 import _lib as ML
from myriad_b200.config import Config, HParams, NLPSolverType
from myriad_b200.engine import Engine

from collections import OrderedDict

MAX_ENGINES = 32   # engines own device workspaces: least-recently-used ones are dropped (solve_with_params in a loop)
_ENGINES: "OrderedDict[tuple, Engine]" = OrderedDict()


def _engine_for(tr) -> Engine:
  d = tr.desc()
  key = (d.system_id, d.optimizer, d.integration_method, d.intervals, d.controls_per_interval, d.T, tuple(d.params[:d.n_params]),
         d.theta, tuple(d.node_hidden[:d.node_num_hidden]))
  eng = _ENGINES.pop(key, None)
  if eng is None:
    eng = Engine(d)
  _ENGINES[key] = eng
  while len(_ENGINES) > MAX_ENGINES:
    _ENGINES.popitem(last=False)
  return eng


def transcription_from_hparams(hp: HParams, system=None):
  """What get_optimizer(hp, cfg, hp.system()) would transcribe (trajectory_optimizers/__init__.py:12-28)."""
  from myriad_b200 import problems as PR
  from myriad_b200.config import OptimizerType, QuadratureRule
  system = hp.system() if system is None else system
  if hp.optimizer == OptimizerType.SHOOTING:
    optid = PR.SHOOTING
  elif hp.optimizer == OptimizerType.COLLOCATION:
    optid = PR.TRAPEZOIDAL if hp.quadrature_rule == QuadratureRule.TRAPEZOIDAL else PR.HERMITE_SIMPSON
  else:
    raise KeyError(hp.optimizer)
  return PR.Transcription(system, optid, hp.integration_method.name, hp.intervals, hp.controls_per_interval)


def _resolve_transcription(hp: HParams, opt_dict: Dict):
  tr = opt_dict.get('transcription')
  if tr is not None:
    return tr
  for key in ('objective', 'constraints'):   # bound methods of one of our optimizers
    owner = getattr(opt_dict.get(key), '__self__', None)
    if owner is not None and getattr(owner, 'transcription', None) is not None:
      return owner.transcription
  tr = transcription_from_hparams(hp)
  guess = np.asarray(opt_dict['guess'], dtype=np.float64)
  if guess.shape != (tr.nvars,):
    raise ValueError(f"opt_dict['guess'] has shape {guess.shape}; hp describes an NLP with {tr.nvars} variables")
  # the kernels will solve hp's transcription: make sure that IS the problem the caller's callables describe
  r = _engine_for(tr).eval(torch.as_tensor(guess).reshape(1, -1).cuda())
  torch.cuda.synchronize()
  for key, mine in (('objective', r.f[0].cpu().numpy()), ('constraints', r.c[0].cpu().numpy())):
    fn = opt_dict.get(key)
    if fn is None:
      continue
    theirs = np.asarray(fn(guess), dtype=np.float64).reshape(np.shape(mine))
    if not np.allclose(theirs, mine, rtol=1e-8, atol=1e-10):
      raise ValueError(f"opt_dict['{key}'] is not the {hp.optimizer.name} transcription of {hp.system.name} that hp describes "
                       "(custom or parametrized callables?).  The CUDA solver evaluates generated device code, not Python "
                       "callables: pass opt_dict['transcription'] (e.g. optimizer.transcription of a parametrized system).")
  return tr


def solve_batch(hp: HParams, cfg: Config, tr, z0: torch.Tensor, lb: torch.Tensor, ub: torch.Tensor) -> Dict[str, torch.Tensor]:
  """Batched solve on the current CUDA device; returns device tensors (z, lam, obj, status, iters, ...)."""
  if hp.nlpsolver == NLPSolverType.EXTRAGRADIENT:
    from myriad_b200.nlp_solvers.extra_gradient import extra_gradient_batch
    return extra_gradient_batch(hp, _engine_for(tr), z0, lb, ub, max_iter=hp.max_iter)
  if hp.nlpsolver != NLPSolverType.IPOPT:
    if hp.nlpsolver in (NLPSolverType.SLSQP, NLPSolverType.TRUST):
      raise NotImplementedError(f"{hp.nlpsolver} is a CPU solver of the reference; the B200 engine serves NLPSolverType.IPOPT "
                                "(interior point) and EXTRAGRADIENT.  See DESIGN.md 'out of scope'.")
    print("Unknown NLP solver. Please choose among", list(NLPSolverType.__members__.keys()))
    raise ValueError
  eng = _engine_for(tr)
  return eng.ipm_solve(z0, lb, ub, max_iter=hp.max_iter)


def solve(hp: HParams, cfg: Config, opt_dict: Dict) -> Dict[str, np.ndarray]:
  """Single-problem solve with the reference's calling convention and result keys."""
  _t1 = time.time()
  tr = _resolve_transcription(hp, opt_dict)
  dev = torch.device("cuda")
  as_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).reshape(1, -1).to(dev).contiguous()
  bounds = np.asarray(opt_dict['bounds'], dtype=np.float64)
  out = solve_batch(hp, cfg, tr, as_dev(opt_dict['guess']), as_dev(bounds[:, 0]), as_dev(bounds[:, 1]))
  z = out["z"][0].cpu().numpy()
  status = int(out["status"][0])
  _t2 = time.time()
  x, u = opt_dict['unravel'](z)
  if cfg.verbose:
    print('Solver exited with success:', status == 0, f"({ML.STATUS_NAMES.get(status, status)}, {int(out['iters'][0])} iterations)")
    print(f'Completed in {_t2 - _t1} seconds.')
    from myriad_b200.utils import get_state_trajectory_and_cost
    system = hp.system()
    opt_x, c = get_state_trajectory_and_cost(hp, system, system.x_0, u)
    print('Cost given by solver:', float(out["obj"][0]))
    print("Cost given by integrating the control trajectory:", c)
    if system.x_T is not None:
      defect = [opt_x[-1][i] - el for i, el in enumerate(system.x_T) if el is not None]
      print("Defect:", defect)
  results = {'x': x, 'u': u, 'xs_and_us': z, 'cost': float(out["obj"][0]), 'lambda': out["lam"][0].cpu().numpy(),
             'status': status, 'iters': int(out["iters"][0]), 'kkt_error': float(out["kkt_err"][0])}
  return results
