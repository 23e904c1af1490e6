"""Hyper-parameters and enums: the reference's configuration surface (myriad/config.py:12-129) without jax.

This file is a schema the north star says to keep ("keeps the OptimizerType / SystemType plugin surface and run.py
entry"): enum members, field names, order and defaults ARE the reference's, on purpose, so that ``run.py`` flags and
``HParams(...)`` constructions carry over unchanged.  Differences: ``hp.key`` is the integer seed (the hot path never
draws JAX random numbers: SURVEY.md section 9-12) and ``batch`` is the one field the B200 engine adds.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Tuple

from myriad_b200.systems import SystemType


class OptimizerType(Enum):
  """Optimizing strategy used to solve the OCP (myriad/config.py:12-17)"""
  COLLOCATION = "COLLOCATION"
  SHOOTING = "SHOOTING"
  FBSM = "FBSM"


class SamplingApproach(Enum):
  """myriad/config.py:20-25"""
  UNIFORM = "UNIFORM"
  TRUE_OPTIMAL = "TRUE_OPTIMAL"
  RANDOM_WALK = "RANDOM_WALK"
  CURRENT_OPTIMAL = "CURRENT_OPTIMAL"


class NLPSolverType(Enum):
  """myriad/config.py:39-44.  IPOPT is served by the batched CUDA interior-point solver (myr_ipm_solve), EXTRAGRADIENT
  by the K1-based primal-dual iteration in myriad_b200/nlp_solvers/extra_gradient.py."""
  SLSQP = "SLSQP"
  TRUST = "TRUST"
  IPOPT = "IPOPT"
  EXTRAGRADIENT = "EXTRAGRADIENT"


class IntegrationMethod(Enum):
  """myriad/config.py:46-50 (the value strings are the reference's)"""
  EULER = "CONSTANT"
  HEUN = "LINEAR"
  MIDPOINT = "MIDPOINT"
  RK4 = "RK4"


class QuadratureRule(Enum):
  """myriad/config.py:53-56"""
  TRAPEZOIDAL = "TRAPEZOIDAL"
  HERMITE_SIMPSON = "HERMITE_SIMPSON"


@dataclass(eq=True, frozen=False)
class HParams:
  """The hyperparameters of the experiment (myriad/config.py:61-112).  The hot path reads the first ten."""
  seed: int = 2019
  system: SystemType = SystemType.CANCERTREATMENT
  optimizer: OptimizerType = OptimizerType.SHOOTING
  nlpsolver: NLPSolverType = NLPSolverType.IPOPT
  integration_method: IntegrationMethod = IntegrationMethod.HEUN
  quadrature_rule: QuadratureRule = QuadratureRule.TRAPEZOIDAL

  max_iter: int = 1000               # maxiter of the NLP solver
  intervals: int = 1                 # COLLOCATION and SHOOTING
  controls_per_interval: int = 100   # SHOOTING
  fbsm_intervals: int = 1000         # FBSM

  # system-identification / neural-ODE experiment knobs: carried for flag compatibility
  sampling_approach: SamplingApproach = SamplingApproach.RANDOM_WALK
  train_size: int = 100
  val_size: int = 3
  test_size: int = 3
  sample_spread: float = 0.05
  start_spread: float = 0.1
  noise_level: float = 0.0
  to_smooth: bool = False
  learning_rate: float = 0.001
  minibatch_size: int = 16
  num_epochs: int = 10_001
  num_experiments: int = 1
  loss_recording_frequency: int = 10
  plot_progress_frequency: int = 10
  early_stop_threshold: int = 30
  early_stop_check_frequency: int = 20
  hidden_layers: Tuple[int, int] = (50, 50)
  num_unrolled: int = 5
  eta_x: float = 1e-1
  eta_lmbda: float = 1e-3
  adam_lr: float = 1e-4

  # addition of the B200 engine: problem instances per launch (instance 0 = the system's own x_0, the others are start
  # states perturbed with start_spread)
  batch: int = 1

  def __post_init__(self):
    """myriad/config.py:97-112: runs once; later mutation leaves the derived fields stale exactly like there."""
    if self.optimizer == OptimizerType.COLLOCATION:
      self.controls_per_interval = 1
    if self.nlpsolver == NLPSolverType.EXTRAGRADIENT:
      self.max_iter *= 10
    system = self.system()
    self.num_steps = self.intervals * self.controls_per_interval
    self.stepsize = system.T / self.num_steps
    self.key = self.seed
    self.state_size = system.x_0.shape[0]
    self.control_size = system.bounds.shape[0] - self.state_size
    self.minibatch_size = min([self.minibatch_size, self.train_size, self.val_size, self.test_size])


@dataclass(eq=True, frozen=False)
class Config:
  """Secondary configurations that should not change experiment results (myriad/config.py:115-129)."""
  verbose: bool = True
  jit: bool = True
  plot: bool = True
  pretty_plotting: bool = True
  load_params_if_saved: bool = True
  figsize: Tuple[float, float] = (8, 6)
  file_extension: str = "png"
