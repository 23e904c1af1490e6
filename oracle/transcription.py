"""CPU oracle: integrators and the three direct transcriptions, restated from the reference.

TEST INFRASTRUCTURE ONLY (see oracle/systems.py header).  Follows

* integrators            myriad/utils.py:22-134
* multiple shooting      myriad/trajectory_optimizers/shooting.py:26-275
* trapezoidal collocation myriad/trajectory_optimizers/collocation/trapezoidal.py:25-192
* Hermite-Simpson        myriad/trajectory_optimizers/collocation/hermite_simpson.py:28-335
* post-solve rollout     myriad/utils.py:258-324

All callables broadcast over leading batch dims of the decision vector ``z[..., nvars]`` so
first derivatives can be taken by one batched complex-step evaluation (oracle/derivatives.py).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .systems import OracleSystem, _xp, torch

EULER, HEUN, MIDPOINT, RK4 = "EULER", "HEUN", "MIDPOINT", "RK4"
METHODS = (EULER, HEUN, MIDPOINT, RK4)


def _cat(parts, like):
  xp = _xp(like)
  return np.concatenate(parts, axis=-1) if xp is np else torch.cat(parts, dim=-1)


def _stack_axis(parts, like, axis):
  xp = _xp(like)
  return np.stack(parts, axis=axis) if xp is np else torch.stack(parts, dim=axis)


def integrate(dynamics_t: Callable, x_0, interval_us, h: float, N: int, ts, method: str):
  """myriad/utils.py:22-73.  x_0[..., d]; interval_us[..., L, m]; ts[..., >=N] (or 1-D).
  Out-of-range control indices clamp to the last entry like JAX does.
  Returns (x_N, states[..., N+1, d])."""
  L = interval_us.shape[-2]

  def U(i):
    return interval_us[..., min(i, L - 1), :]

  def Tm(i):
    return ts[..., min(i, ts.shape[-1] - 1)]

  x = x_0
  out = [x]
  for idx in range(N):
    t = Tm(idx)
    if method == EULER:  # :53-54
      x = x + h * dynamics_t(x, U(idx), t)
    elif method == HEUN:  # :41-44
      k1 = dynamics_t(x, U(idx), t)
      k2 = dynamics_t(x + h * k1, U(idx + 1), t + h)
      x = x + h / 2 * (k1 + k2)
    elif method == MIDPOINT:  # :47-50 (full Euler step to the "midpoint" state)
      x_mid = x + h * dynamics_t(x, U(idx), t)
      u_mid = (U(idx) + U(idx + 1)) / 2
      x = x + h * dynamics_t(x_mid, u_mid, t + h / 2)
    elif method == RK4:  # :33-38, :64-65
      u1, u2, u3 = U(2 * idx), U(2 * idx + 1), U(2 * idx + 2)
      k1 = dynamics_t(x, u1, t)
      k2 = dynamics_t(x + h * k1 / 2, u2, t + h / 2)
      k3 = dynamics_t(x + h * k2 / 2, u2, t + h / 2)
      k4 = dynamics_t(x + h * k3, u3, t + h)
      x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    else:
      raise KeyError(method)
    out.append(x)
  return x, _stack_axis(out, x_0, -2)


def integrate_time_independent(dynamics: Callable, x_0, interval_us, h: float, N: int, method: str):
  """myriad/utils.py:80-131"""
  zero_t = np.zeros(1)
  return integrate(lambda x, u, t: dynamics(x, u), x_0, interval_us, h, N, zero_t, method)


class Transcription:
  """Common part of the oracle's NLPs: mirrors the TrajectoryOptimizer fields
  (myriad/trajectory_optimizers/base.py:33-52): guess, bounds, objective, constraints, unravel."""
  system: OracleSystem
  nx_nodes: int
  nu_nodes: int

  @property
  def n(self):
    return self.system.n

  @property
  def m(self):
    return self.system.m

  @property
  def nvars(self):
    return self.nx_nodes * self.n + self.nu_nodes * self.m

  def unravel(self, z):
    """ravel_pytree((x, u)): states time-major then controls time-major."""
    nx = self.nx_nodes * self.n
    x = z[..., :nx].reshape(z.shape[:-1] + (self.nx_nodes, self.n))
    u = z[..., nx:].reshape(z.shape[:-1] + (self.nu_nodes, self.m))
    return x, u

  def _control_bounds(self):
    # control-major fill (shooting.py:264-267, trapezoidal.py:58-61, hermite_simpson.py:71-74);
    # identical to time-major only for m == 1 -- replicated as coded.
    m, L = self.m, self.nu_nodes
    ub = np.empty((L * m, 2))
    for i in range(m, 0, -1):
      ub[(m - i) * L:(m - i + 1) * L] = self.system.bounds[-i]
    return ub


def _state_guess_rows(system: OracleSystem, n_nodes: int, rollout: Callable[[], np.ndarray]):
  """Per-row linspace / rollout mix (shooting.py:55-73, trapezoidal.py:36-50)."""
  if system.x_T is not None:
    rows = []
    rolled = None
    for i in range(len(system.x_T)):
      if system.x_T[i] is not None:
        rows.append(np.linspace(system.x_0[i], system.x_T[i], num=n_nodes))
      else:
        if rolled is None:
          rolled = rollout()
        rows.append(rolled[:, i])
    return np.stack(rows, axis=1)
  return rollout()


class Shooting(Transcription):
  """myriad/trajectory_optimizers/shooting.py:15-278"""

  def __init__(self, system: OracleSystem, intervals: int, controls_per_interval: int, method: str = HEUN):
    self.system = system
    self.K = intervals
    self.cpi = controls_per_interval
    self.method = method
    self.num_steps = intervals * controls_per_interval  # :26
    self.step_size = system.T / self.num_steps  # :27
    self.interval_size = system.T / intervals  # :28
    self.mc = 2 if method == RK4 else 1  # :31
    self.nx_nodes = intervals + 1
    self.nu_nodes = self.mc * self.num_steps + 1
    n, m = system.n, system.m

    u_guess = np.zeros((self.nu_nodes, m))  # :47

    def rollout():
      _, xs = integrate_time_independent(system.dynamics, system.x_0, u_guess[::self.mc * self.cpi],
                                         self.interval_size, intervals, method)  # :64-67, :72-74
      return xs

    x_guess = _state_guess_rows(system, intervals + 1, rollout)
    self.x_guess, self.u_guess = x_guess, u_guess
    self.guess = np.concatenate([x_guess.ravel(), u_guess.ravel()])  # :75

    xb = np.zeros((intervals + 1, n, 2))  # :248-258
    xb[:, :, :] = system.bounds[:-m]
    xb[0, :, :] = system.x_0[:, None]
    if system.x_T is not None:
      for i in range(len(system.x_T)):
        if system.x_T[i] is not None:
          xb[-1, i, :] = system.x_T[i]
    self.x_bounds = xb.reshape(-1, 2)
    self.u_bounds = self._control_bounds()
    self.bounds = np.vstack([self.x_bounds, self.u_bounds])
    self.ncon = intervals * n

  def reorganize_controls(self, us):
    """:100-130 -> [..., K, mc*cpi+1, m]"""
    M = self.mc * self.cpi
    lead = us.shape[:-2]
    body = us[..., :-1, :].reshape(lead + (self.K, M, self.m))
    nxt = us[..., ::M, :][..., 1:, :][..., None, :]
    xp = _xp(us)
    return np.concatenate([body, nxt], axis=-2) if xp is np else torch.cat([body, nxt], dim=-2)

  def reorganize_times(self, ts):
    """:132-142 (ignores mc)"""
    body = ts[:-1].reshape(self.K, self.cpi)
    nxt = ts[::self.cpi][1:][:, None]
    return np.concatenate([body, nxt], axis=1)

  def objective(self, z):
    """:169-210"""
    xs, us = self.unravel(z)
    sysm = self.system
    ctrl = self.reorganize_controls(us)
    t = self.reorganize_times(np.linspace(0., sysm.T, num=self.num_steps + 1))
    start = xs[..., :-1, :]
    zeros = start[..., :1] * 0
    aug0 = _cat([start, zeros], z)

    def aug_dyn(xc, u, tt):  # :80-92
      x = xc[..., :-1]
      f = sysm.dynamics(x, u)
      g = sysm.cost(x, u, tt)
      return _cat([f, g[..., None]], z)

    end, _ = integrate(aug_dyn, aug0, ctrl, self.step_size, self.cpi, t, self.method)
    costs = end[..., -1].sum(-1)
    if sysm.terminal_cost:  # :205-208
      costs = costs + sysm.terminal_cost_fn(end[..., -1, :-1], us[..., -1, :])
    return costs

  def constraints(self, z):
    """:230-241  ravel(px - xs[1:])"""
    xs, us = self.unravel(z)
    px, _ = integrate_time_independent(self.system.dynamics, xs[..., :-1, :], self.reorganize_controls(us),
                                       self.step_size, self.cpi, self.method)
    d = px - xs[..., 1:, :]
    return d.reshape(d.shape[:-2] + (-1,))


class Trapezoid(Transcription):
  """myriad/trajectory_optimizers/collocation/trapezoidal.py:15-209"""

  def __init__(self, system: OracleSystem, intervals: int, method: str = HEUN):
    self.system = system
    self.N = intervals
    self.h = system.T / intervals  # :26
    self.method = method
    n, m = system.n, system.m
    self.nx_nodes = self.nu_nodes = intervals + 1
    u_guess = np.zeros((intervals + 1, m))  # :34

    def rollout():
      _, xs = integrate_time_independent(system.dynamics, system.x_0, u_guess, self.h, intervals, method)
      return xs

    x_guess = _state_guess_rows(system, intervals + 1, rollout)  # :36-50
    self.x_guess, self.u_guess = x_guess, u_guess
    self.guess = np.concatenate([x_guess.ravel(), u_guess.ravel()])

    xb = np.empty((intervals + 1, n, 2))  # :66-71
    xb[:, :, :] = system.bounds[:-m]
    xb[0, :, :] = system.x_0[:, None]
    if system.x_T is not None:
      xb[-m, :, :] = np.asarray(system.x_T, dtype=np.float64)[:, None]  # node index -m
    self.x_bounds = xb.reshape(-1, 2)
    self.u_bounds = self._control_bounds()
    self.bounds = np.vstack([self.x_bounds, self.u_bounds])
    self.ncon = intervals * n

  def objective(self, z):
    """:115-128"""
    x, u = self.unravel(z)
    sysm = self.system
    t = np.linspace(0, sysm.T, num=self.N + 1)
    g = sysm.cost(x, u, t)
    cost = ((self.h / 2) * (g[..., :-1] + g[..., 1:])).sum(-1)
    if sysm.terminal_cost:
      cost = cost + sysm.terminal_cost_fn(x[..., -1, :], u[..., -1, :])
    return cost

  def constraints(self, z):
    """:151-163, :183-192   c_k = h/2 (f_k + f_{k+1}) - (x_{k+1} - x_k)"""
    x, u = self.unravel(z)
    f = self.system.dynamics(x, u)
    d = (self.h / 2) * (f[..., :-1, :] + f[..., 1:, :]) - (x[..., 1:, :] - x[..., :-1, :])
    return d.reshape(d.shape[:-2] + (-1,))


class HermiteSimpson(Transcription):
  """myriad/trajectory_optimizers/collocation/hermite_simpson.py:15-351"""

  def __init__(self, system: OracleSystem, intervals: int, method: str = HEUN):
    self.system = system
    self.N = intervals
    self.h = system.T / intervals  # :28
    self.method = method
    n, m = system.n, system.m
    self.nx_nodes = self.nu_nodes = 2 * intervals + 1
    u_guess = np.zeros((2 * intervals + 1, m))  # :37
    if system.x_T is not None:  # :40-43
      x_guess = np.linspace(system.x_0, np.asarray(system.x_T, dtype=np.float64), num=2 * intervals + 1)
    else:
      x_guess = np.ones((2 * intervals + 1, n)) * 0.1
    self.x_guess, self.u_guess = x_guess, u_guess
    self.guess = np.concatenate([x_guess.ravel(), u_guess.ravel()])

    xb = np.zeros((2 * intervals + 1, n, 2))  # :55-65
    xb[:, :, :] = system.bounds[:-m]
    xb[0, :, :] = system.x_0[:, None]
    if system.x_T is not None:
      for i in range(len(system.x_T)):
        if system.x_T[i] is not None:
          xb[-1, i, :] = system.x_T[i]
    self.x_bounds = xb.reshape(-1, 2)
    self.u_bounds = self._control_bounds()
    self.bounds = np.vstack([self.x_bounds, self.u_bounds])
    self.ncon = 2 * intervals * n

  def objective(self, z):
    """:243-257   sum h/6 (g_k + 4 g_m + g_{k+1}); no terminal-cost term"""
    x, u = self.unravel(z)
    t = np.linspace(0, self.system.T, num=2 * self.N + 1)
    g = self.system.cost(x, u, t)
    return ((self.h / 6) * (g[..., 0:-1:2] + 4 * g[..., 1::2] + g[..., 2::2])).sum(-1)

  def constraints(self, z):
    """:110-128 defects, :153-170 interpolation, :325-335 hstack(defects, interpolations)"""
    x, u = self.unravel(z)
    f = self.system.dynamics(x, u)
    xs, xm, xe = x[..., 0:-1:2, :], x[..., 1::2, :], x[..., 2::2, :]
    fs, fm, fe = f[..., 0:-1:2, :], f[..., 1::2, :], f[..., 2::2, :]
    defect = (xe - xs) - (self.h / 6) * (fs + 4 * fm + fe)
    interp = xm - 0.5 * (xs + xe) - (self.h / 8) * (fs - fe)
    lead = defect.shape[:-2]
    return _cat([defect.reshape(lead + (-1,)), interp.reshape(lead + (-1,))], z)


def get_state_trajectory_and_cost(system: OracleSystem, intervals: int, cpi: int, method: str, start_state, us):
  """myriad/utils.py:258-298.  us[..., L, m] -> (states[..., num_steps+1, n], cost[...])"""
  num_steps = intervals * cpi
  step = system.T / num_steps
  times = np.linspace(0., system.T, num=num_steps + 1)

  def aug_dyn(xc, u, t):
    x = xc[..., :-1]
    return _cat([system.dynamics(x, u), system.cost(x, u, t)[..., None]], xc)

  start = _cat([start_state, start_state[..., :1] * 0], start_state)
  _, sc = integrate(aug_dyn, start, us, step, num_steps, times, method)
  states = sc[..., :-1]
  cost = sc[..., -1, -1]
  if system.terminal_cost:
    cost = cost + system.terminal_cost_fn(sc[..., -1, :-1], us[..., -1, :])
  return states, cost


def get_defect(system: OracleSystem, xs) -> Optional[np.ndarray]:
  """myriad/utils.py:313-324"""
  if system.x_T is None:
    return None
  idx = [i for i in range(len(system.x_T)) if system.x_T[i] is not None]
  tgt = np.array([system.x_T[i] for i in idx], dtype=np.float64)
  return xs[..., -1, idx] - tgt


def make_transcription(system: OracleSystem, optimizer: str, intervals: int, cpi: int = 1,
                       method: str = HEUN, quadrature: str = "TRAPEZOIDAL") -> Transcription:
  """get_optimizer (myriad/trajectory_optimizers/__init__.py:12-28)"""
  if optimizer == "COLLOCATION":
    if quadrature == "TRAPEZOIDAL":
      return Trapezoid(system, intervals, method)
    if quadrature == "HERMITE_SIMPSON":
      return HermiteSimpson(system, intervals, method)
    raise KeyError(quadrature)
  if optimizer == "SHOOTING":
    return Shooting(system, intervals, cpi, method)
  raise KeyError(optimizer)
