"""NLP solve: mirror of myriad/nlp_solvers/__init__.py:18-98 on top of the batched CUDA interior-point
solver (myr_ipm_solve), which replaces the cyipopt call at :56-58.

``solve(hp, cfg, opt_dict)`` keeps the reference's signature and result dictionary
(``x, u, xs_and_us, cost, lambda``, :90-96).  ``opt_dict`` additionally carries the optimizer's
``Transcription`` (key ``'transcription'``): objective/constraint *callables* cannot be traced into CUDA the way
jax traces them, so the kernels are selected by the transcription's descriptor instead.
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np
import torch

from myriad_b200 import _lib as ML
from myriad_b200.config import Config, HParams, NLPSolverType
from myriad_b200.engine import Engine

_ENGINES: Dict[tuple, Engine] = {}


def _engine_for(tr) -> Engine:
  d = tr.desc()
  key = (d.system_id, d.optimizer, d.integration_method, d.intervals, d.controls_per_interval, d.T, tuple(d.params[:d.n_params]),
         d.theta, tuple(d.node_hidden[:d.node_num_hidden]))
  if key not in _ENGINES:
    _ENGINES[key] = Engine(d)
  return _ENGINES[key]


def solve_batch(hp: HParams, cfg: Config, tr, z0: torch.Tensor, lb: torch.Tensor, ub: torch.Tensor) -> Dict[str, torch.Tensor]:
  """Batched solve on the current CUDA device; returns device tensors (z, lam, obj, status, iters, ...)."""
  if hp.nlpsolver != NLPSolverType.IPOPT:
    if hp.nlpsolver in (NLPSolverType.SLSQP, NLPSolverType.TRUST, NLPSolverType.EXTRAGRADIENT):
      raise NotImplementedError(f"{hp.nlpsolver} is a CPU solver of the reference; the B200 engine serves NLPSolverType.IPOPT "
                                "(interior point).  See DESIGN.md 'out of scope'.")
    print("Unknown NLP solver. Please choose among", list(NLPSolverType.__members__.keys()))
    raise ValueError
  eng = _engine_for(tr)
  return eng.ipm_solve(z0, lb, ub, max_iter=hp.max_iter)


def solve(hp: HParams, cfg: Config, opt_dict: Dict) -> Dict[str, np.ndarray]:
  """Single-problem solve with the reference's calling convention and result keys."""
  _t1 = time.time()
  tr = opt_dict['transcription']
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
