"""Post-solve helpers: mirror of myriad/utils.py:258-324 on top of the CUDA rollout kernel."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from myriad_b200 import problems as PR
from myriad_b200.config import HParams, OptimizerType
from myriad_b200.engine import Engine


def _rollout_engine(hp: HParams, system) -> Engine:
  optid = PR.SHOOTING if hp.optimizer == OptimizerType.SHOOTING else PR.TRAPEZOIDAL
  tr = PR.Transcription(system, optid, hp.integration_method.name, hp.intervals, hp.controls_per_interval)
  eng = Engine.__new__(Engine)
  eng.desc, eng._ws = tr.desc(), None
  return eng


def get_state_trajectory_and_cost_batch(hp: HParams, system, start_states: torch.Tensor, us: torch.Tensor):
  """Batched myriad/utils.py:258-298: integrate the (true) system under controls ``us`` [B, rows, m] from
  ``start_states`` [B, n] with hp.integration_method over intervals * controls_per_interval steps."""
  eng = _rollout_engine(hp, system)
  return eng.rollout_cost(us.contiguous(), start_states.contiguous(), want_states=True)


def get_state_trajectory_and_cost(hp: HParams, system, start_state, us) -> Tuple[np.ndarray, float]:
  x0 = torch.as_tensor(np.asarray(start_state, dtype=np.float64)).reshape(1, -1).cuda()
  u = torch.as_tensor(np.asarray(us, dtype=np.float64)).reshape(1, np.shape(us)[0], -1).cuda()
  xs, cost = get_state_trajectory_and_cost_batch(hp, system, x0, u)
  return xs[0].cpu().numpy(), float(cost[0])


def get_defect(system, learned_xs) -> Optional[np.ndarray]:
  """myriad/utils.py:313-324"""
  defect = None
  if system.x_T is not None:
    defect = []
    for i, s in enumerate(learned_xs[-1]):
      if system.x_T[i] is not None:
        defect.append(s - system.x_T[i])
  if defect is not None:
    defect = np.array(defect)
  return defect
