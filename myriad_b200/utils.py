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


def generate_dataset(hp: HParams, cfg, given_us=None, seed: Optional[int] = None) -> np.ndarray:
  """Mirror of myriad/utils.py:327-444: ``train_size + val_size + test_size`` random control trajectories and the state
  trajectories the TRUE system follows under them -- every rollout of the dataset in ONE launch of the CUDA rollout
  kernel (the reference vmaps integrate_time_independent over the trajectories, :422-424).

  Control sampling follows the reference draw for draw where it uses NumPy's global generator (RANDOM_WALK, :349-358:
  seed it with np.random.seed(hp.seed) as useful_scripts.py does); where the reference draws from jax.random (UNIFORM,
  TRUE_OPTIMAL / CURRENT_OPTIMAL, start-state spread, observation noise) a torch CPU generator seeded with ``seed``
  (default hp.seed) is used instead -- JAX's PRNG stream is not reproducible without jax.
  Returns xs_and_us [total, num_steps + 1, n + m] as a NumPy array."""
  from myriad_b200.config import SamplingApproach
  system = hp.system()
  total = hp.train_size + hp.val_size + hp.test_size
  n, m = hp.state_size, hp.control_size
  b = np.asarray(system.bounds, dtype=np.float64)
  x_lower, x_upper, u_lower, u_upper = b[:n, 0], b[:n, 1], b[n:, 0], b[n:, 1]
  if np.isinf(u_lower).any() or np.isinf(u_upper).any():
    raise Exception("infinite control bounds, aborting")
  if np.isinf(x_lower).any() or np.isinf(x_upper).any():
    raise Exception("infinite state bounds, aborting")
  gen = torch.Generator(device="cpu").manual_seed(int(hp.seed if seed is None else seed))
  spread = (u_upper - u_lower) * hp.sample_spread
  if hp.sampling_approach == SamplingApproach.RANDOM_WALK:
    all_us = np.random.uniform(u_lower, u_upper, (total, 1, m))
    for _ in range(hp.num_steps):
      next_us = np.random.normal(0, spread, (total, 1, m))
      all_us = np.concatenate((all_us, np.clip(next_us + all_us[:, -1:, :], u_lower, u_upper)), axis=1)
  elif hp.sampling_approach == SamplingApproach.UNIFORM or given_us is None:
    r = torch.rand(total, hp.num_steps + 1, m, generator=gen, dtype=torch.float64).numpy()
    all_us = (u_lower + r * (u_upper - u_lower)) * 0.75
  elif hp.sampling_approach in (SamplingApproach.TRUE_OPTIMAL, SamplingApproach.CURRENT_OPTIMAL):
    noise = torch.randn(total, hp.num_steps + 1, m, generator=gen, dtype=torch.float64).numpy() * (u_upper - u_lower) * hp.sample_spread
    all_us = np.clip(np.asarray(given_us, dtype=np.float64).reshape(1, hp.num_steps + 1, m) + noise, u_lower, u_upper)
  else:
    raise Exception("Unknown sampling approach, please choose among", [s.name for s in SamplingApproach])
  start = np.repeat(np.asarray(system.x_0, dtype=np.float64)[None], total, axis=0)
  if hp.start_spread > 0.:
    start = np.clip(start + torch.randn(total, n, generator=gen, dtype=torch.float64).numpy() * hp.start_spread, x_lower, x_upper)
  xs, _ = get_state_trajectory_and_cost_batch(hp, system, torch.as_tensor(start).cuda(), torch.as_tensor(all_us).cuda())
  all_xs = xs.cpu().numpy()
  if hp.noise_level > 0.:
    all_xs = all_xs + torch.randn(all_xs.shape, generator=gen, dtype=torch.float64).numpy() * (x_upper - x_lower) * hp.noise_level
  all_xs = np.clip(all_xs, x_lower, x_upper)
  out = np.concatenate((all_xs, all_us), axis=2)
  assert np.isfinite(out).all()
  return out
