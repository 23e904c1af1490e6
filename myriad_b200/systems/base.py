"""System descriptors: the B200 engine's mirror of myriad/systems/base.py:11-111.

A system here is DATA (x_0, x_T, T, bounds, parameters) plus the name of its generated device code
(csrc/systems_gen.cuh, produced by tools/gen_systems.py from the same formulas).  ``dynamics`` / ``cost``
are therefore not Python callables that the solver traces, as in the reference: they are evaluated on the
GPU by the kernels.  The Python methods below evaluate one point through the same kernels so user code that
calls ``system.dynamics(x, u)`` keeps working.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np


@dataclass
class FiniteHorizonControlSystem(object):
  x_0: np.ndarray
  """State at time 0"""
  x_T: Optional[Sequence]
  """State at time T (entries may be None)"""
  T: float
  """Duration of trajectory"""
  bounds: np.ndarray
  """State and control bounds (states first)"""
  terminal_cost: bool = False
  discrete: bool = False
  device_name: str = ""
  """SystemType name of the generated device implementation"""
  params: Sequence[float] = field(default_factory=list)
  """parameter vector in the order of tools/gen_systems.py"""

  @property
  def state_size(self) -> int:
    return int(np.asarray(self.x_0).shape[0])

  @property
  def control_size(self) -> int:
    return int(np.asarray(self.bounds).shape[0]) - self.state_size

  def terminal_cost_fn(self, x_T, u_T, T=None):
    return 0
