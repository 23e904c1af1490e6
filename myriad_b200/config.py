"""Hyper-parameters and enums: the reference's configuration surface (myriad/config.py:12-129) without jax.

Field names, order, defaults and the derived fields of ``HParams.__post_init__`` follow the reference so that ``run.py``
flags and ``HParams(...)`` constructions carry over; ``hp.key`` is the integer seed (the hot path never draws JAX random
numbers: SURVEY.md section 9-12).  The two dataclasses are generated from the field tables below; ``batch`` is the one
field the B200 engine adds.
"""
from __future__ import annotations

import dataclasses
from enum import Enum
from typing import Tuple

from myriad_b200.systems import SystemType


def _enum(name: str, members: str, doc: str = "") -> type:
  """Enum whose member values are given as ``NAME=value`` pairs (value defaults to the name)."""
  pairs = [m.split("=") if "=" in m else (m, m) for m in members.split()]
  e = Enum(name, [(k, v) for k, v in pairs], module=__name__)
  e.__doc__ = doc
  return e


# myriad/config.py:12-17, 20-25, 39-44, 46-50, 53-56 (IntegrationMethod keeps the reference's value strings)
OptimizerType = _enum("OptimizerType", "COLLOCATION SHOOTING FBSM", "Optimizing strategy used to solve the OCP")
SamplingApproach = _enum("SamplingApproach", "UNIFORM TRUE_OPTIMAL RANDOM_WALK CURRENT_OPTIMAL")
NLPSolverType = _enum("NLPSolverType", "SLSQP TRUST IPOPT EXTRAGRADIENT",
                      "IPOPT is served by the batched CUDA interior-point solver (myr_ipm_solve)")
IntegrationMethod = _enum("IntegrationMethod", "EULER=CONSTANT HEUN=LINEAR MIDPOINT=MIDPOINT RK4=RK4")
QuadratureRule = _enum("QuadratureRule", "TRAPEZOIDAL HERMITE_SIMPSON")

# (name, type, default) in the reference's order (myriad/config.py:63-95); the hot path reads the first ten
_HPARAM_FIELDS = [
  ("seed", int, 2019), ("system", SystemType, SystemType.CANCERTREATMENT), ("optimizer", OptimizerType, OptimizerType.SHOOTING),
  ("nlpsolver", NLPSolverType, NLPSolverType.IPOPT), ("integration_method", IntegrationMethod, IntegrationMethod.HEUN),
  ("quadrature_rule", QuadratureRule, QuadratureRule.TRAPEZOIDAL),
  ("max_iter", int, 1000), ("intervals", int, 1), ("controls_per_interval", int, 100), ("fbsm_intervals", int, 1000),
  # system-identification / neural-ODE experiment knobs: carried for flag compatibility, unused by the B200 hot path
  ("sampling_approach", SamplingApproach, SamplingApproach.RANDOM_WALK), ("train_size", int, 100), ("val_size", int, 3),
  ("test_size", int, 3), ("sample_spread", float, 0.05), ("start_spread", float, 0.1), ("noise_level", float, 0.0),
  ("to_smooth", bool, False), ("learning_rate", float, 0.001), ("minibatch_size", int, 16), ("num_epochs", int, 10_001),
  ("num_experiments", int, 1), ("loss_recording_frequency", int, 10), ("plot_progress_frequency", int, 10),
  ("early_stop_threshold", int, 30), ("early_stop_check_frequency", int, 20), ("hidden_layers", Tuple[int, int], (50, 50)),
  ("num_unrolled", int, 5), ("eta_x", float, 1e-1), ("eta_lmbda", float, 1e-3), ("adam_lr", float, 1e-4),
  # addition of the B200 engine: problem instances per launch (instance 0 = the system's own x_0, the others are start
  # states perturbed with start_spread)
  ("batch", int, 1),
]

# myriad/config.py:115-129
_CONFIG_FIELDS = [("verbose", bool, True), ("jit", bool, True), ("plot", bool, True), ("pretty_plotting", bool, True),
                  ("load_params_if_saved", bool, True), ("figsize", Tuple[float, float], (8, 6)), ("file_extension", str, "png")]


def _derive(self) -> None:
  """HParams.__post_init__ of the reference (myriad/config.py:97-112): runs once, later mutation leaves the derived
  fields stale exactly like there."""
  if self.optimizer == OptimizerType.COLLOCATION:
    self.controls_per_interval = 1                      # :98-99
  if self.nlpsolver == NLPSolverType.EXTRAGRADIENT:
    self.max_iter *= 10                                 # :100-101
  system = self.system()
  self.num_steps = self.intervals * self.controls_per_interval
  self.stepsize = system.T / self.num_steps
  self.key = self.seed
  self.state_size = system.x_0.shape[0]
  self.control_size = system.bounds.shape[0] - self.state_size
  self.minibatch_size = min(self.minibatch_size, self.train_size, self.val_size, self.test_size)  # :112


def _make(name: str, table, doc: str, post=None) -> type:
  ns = {"__doc__": doc}
  if post is not None:
    ns["__post_init__"] = post
  cls = dataclasses.make_dataclass(name, [(n, t, dataclasses.field(default=d)) for n, t, d in table], namespace=ns, eq=True)
  cls.__module__ = __name__
  return cls


HParams = _make("HParams", _HPARAM_FIELDS, "The hyperparameters of the experiment (myriad/config.py:61-112).", _derive)
Config = _make("Config", _CONFIG_FIELDS, "Secondary configurations that should not change experiment results "
                                         "(myriad/config.py:115-129).")
