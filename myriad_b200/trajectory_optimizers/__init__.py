"""get_optimizer factory: mirror of myriad/trajectory_optimizers/__init__.py:12-28."""
from __future__ import annotations

from myriad_b200 import problems as PR
from myriad_b200.config import Config, HParams, OptimizerType, QuadratureRule
from myriad_b200.trajectory_optimizers.base import TrajectoryOptimizer


class TrapezoidalCollocationOptimizer(TrajectoryOptimizer):
  """myriad/trajectory_optimizers/collocation/trapezoidal.py:15-209"""

  def __init__(self, hp: HParams, cfg: Config, system) -> None:
    super().__init__(hp, cfg, system, PR.Transcription(system, PR.TRAPEZOIDAL, hp.integration_method.name, hp.intervals, 1))


class HermiteSimpsonCollocationOptimizer(TrajectoryOptimizer):
  """myriad/trajectory_optimizers/collocation/hermite_simpson.py:15-351"""

  def __init__(self, hp: HParams, cfg: Config, system) -> None:
    super().__init__(hp, cfg, system, PR.Transcription(system, PR.HERMITE_SIMPSON, hp.integration_method.name, hp.intervals, 1))


class MultipleShootingOptimizer(TrajectoryOptimizer):
  """myriad/trajectory_optimizers/shooting.py:15-278"""

  def __init__(self, hp: HParams, cfg: Config, system, key=None) -> None:
    super().__init__(hp, cfg, system, PR.Transcription(system, PR.SHOOTING, hp.integration_method.name, hp.intervals,
                                                       hp.controls_per_interval))


def get_optimizer(hp: HParams, cfg: Config, system) -> TrajectoryOptimizer:
  """Helper function to fetch the desired optimizer for system resolution"""
  if hp.optimizer == OptimizerType.COLLOCATION:
    if hp.quadrature_rule == QuadratureRule.TRAPEZOIDAL:
      optimizer = TrapezoidalCollocationOptimizer(hp, cfg, system)
    elif hp.quadrature_rule == QuadratureRule.HERMITE_SIMPSON:
      optimizer = HermiteSimpsonCollocationOptimizer(hp, cfg, system)
    else:
      raise KeyError
  elif hp.optimizer == OptimizerType.SHOOTING:
    optimizer = MultipleShootingOptimizer(hp, cfg, system)
  elif hp.optimizer == OptimizerType.FBSM:
    from myriad_b200.trajectory_optimizers.fbsm import FBSM
    optimizer = FBSM(hp, cfg, system)
  else:
    raise KeyError
  return optimizer
