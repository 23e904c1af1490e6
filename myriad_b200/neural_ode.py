"""Neural-ODE planning surface: the part of myriad/neural_ode/create_node.py:36-173 the trajectory-optimization hot path
touches -- the MLP definition (net_fn, :110-117), its parameter mapping (:124-131) and load/save -- plus
``plan_with_node_model`` (myriad/utils.py:230-242).  Training (node_training.py), dataset generation and the experiment
glue are outside the B200 hot path ("out of scope").

The network is  Linear(h_1) -> sigmoid -> ... -> Linear(h_k) -> sigmoid -> Linear(n)  on concat(x, u), hk.Linear being
``x @ w + b`` with ``w`` of shape (in, out).  On the device it is evaluated for all collocation nodes of an instance at
once on the fp64 tensor cores (csrc/node_mlp.cuh).
"""
from __future__ import annotations

import math
import pickle as pkl
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

from myriad_b200.config import Config, HParams
from myriad_b200.systems import NodeSystem, mlp_layers


def init_params(n_in: int, hidden, n_out: int, seed: int = 42) -> Dict[str, Dict[str, np.ndarray]]:
  """haiku's default hk.Linear initialisation (truncated normal with stddev 1/sqrt(fan_in), zero bias) drawn with
  NumPy's PCG64 (jax's PRNG is not available here), keys linear, linear_1, ... as create_node.py:124-131 expects."""
  rng = np.random.Generator(np.random.PCG64(seed))
  sizes = [n_in] + list(hidden) + [n_out]
  params = {}
  for i, (fi, fo) in enumerate(zip(sizes[:-1], sizes[1:])):
    w = np.clip(rng.standard_normal((fi, fo)), -2.0, 2.0) / math.sqrt(fi)
    params["linear" if i == 0 else f"linear_{i}"] = {"w": w, "b": np.zeros(fo)}
  return params


@dataclass
class NeuralODE(object):
  """Holds what planning needs from the reference's NeuralODE: hp, cfg, the true system, the MLP parameters and the
  optimizer built over the NODE dynamics (create_node.py:36-56, :110-134)."""
  hp: HParams
  cfg: Config
  params: Optional[Dict] = None
  seed: int = 42

  def __post_init__(self) -> None:
    self.system = self.hp.system()
    self.num_steps = self.hp.intervals * self.hp.controls_per_interval
    self.stepsize = self.system.T / self.num_steps
    if self.params is None:
      n, m = self.system.state_size, self.system.control_size
      self.params = init_params(n + m, self.hp.hidden_layers, n, self.seed)
    self._optimizer = None

  # the reference's net.apply(params, x_and_u), on the host (debugging / plotting; the solver evaluates it on the GPU)
  def apply(self, params, x_and_u: np.ndarray) -> np.ndarray:
    h = np.asarray(x_and_u, dtype=np.float64)
    layers = mlp_layers(params)
    for i, (w, b) in enumerate(layers):
      h = h @ w + b
      if i + 1 < len(layers):
        h = 1.0 / (1.0 + np.exp(-h))
    return h

  def save_params(self, filename: str) -> None:
    pkl.dump(self.params, open(filename, 'wb'))

  def load_params(self, params_pickle: str) -> None:
    """create_node.py:122-134 incl. the renaming of nested haiku keys; .npz files with 'linear/w' keys load too."""
    if params_pickle.endswith(".npz"):
      self.params = dict(np.load(params_pickle))
    else:
      temp = dict(pkl.load(open(params_pickle, 'rb')))
      if 'linear/~/linear' in temp:
        temp['linear_1'] = temp.pop('linear/~/linear')
      if 'linear/~/linear/~/linear' in temp:
        temp['linear_2'] = temp.pop('linear/~/linear/~/linear')
      self.params = temp
    self._optimizer = None

  @property
  def node_system(self) -> NodeSystem:
    return NodeSystem(self, self.system)

  @property
  def optimizer(self):
    """optimizer over the NODE dynamics (guess / bounds are the true system's, as in the reference where the optimizer
    is built before system.dynamics is swapped: myriad/utils.py:231-236)"""
    if self._optimizer is None:
      from myriad_b200.trajectory_optimizers import get_optimizer
      self._optimizer = get_optimizer(self.hp, self.cfg, self.node_system)
    return self._optimizer


def plan_with_node_model(node: NeuralODE) -> Tuple[np.ndarray, np.ndarray]:
  """myriad/utils.py:230-242: solve the planning problem whose dynamics are the NODE's MLP; returns (x, u)."""
  solved_results = node.optimizer.solve()
  return solved_results['x'], solved_results['u']


def plan_with_node_model_batch(node: NeuralODE, x0s):
  """Batched variant (one planning problem per start state in one launch); returns device tensors (x, u, status)."""
  sol = node.optimizer.solve_batch(x0s)
  return sol['x'], sol['u'], sol['status']
