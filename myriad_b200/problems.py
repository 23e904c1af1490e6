"""Batched NLP data (initial guess, bounds) for B problem instances that differ in their start state.

Mirrors, per instance, what the reference optimizers' constructors build:
  shooting     myriad/trajectory_optimizers/shooting.py:47-77 (guess), :248-275 (bounds)
  trapezoid    myriad/trajectory_optimizers/collocation/trapezoidal.py:34-52, :58-78
  Hermite-S.   myriad/trajectory_optimizers/collocation/hermite_simpson.py:37-48, :55-81
All tensors are built on the device with torch ops; state rows that the reference obtains by rolling
out the dynamics under the zero control guess come from the CUDA rollout kernel (myr_rollout_cost).
The reference has no batching over instances (SURVEY.md headline facts): row b of every tensor is
exactly what the reference would build for a system whose x_0 is x0s[b].
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib as ML
from .engine import Engine
from .systems.base import FiniteHorizonControlSystem

SHOOTING, TRAPEZOIDAL, HERMITE_SIMPSON = ML.OPT_SHOOTING, ML.OPT_TRAPEZOIDAL, ML.OPT_HERMITE_SIMPSON


@dataclass
class Transcription:
  system: FiniteHorizonControlSystem
  optimizer: int          # ML.OPT_*
  method: str             # EULER | HEUN | MIDPOINT | RK4
  intervals: int
  cpi: int = 1

  def __post_init__(self):
    if self.optimizer != SHOOTING:
      self.cpi = 1
    self.n = self.system.state_size
    self.m = self.system.control_size
    self.mc = 2 if (self.optimizer == SHOOTING and self.method == "RK4") else 1
    if self.optimizer == SHOOTING:
      self.nx_nodes = self.intervals + 1
      self.nu_nodes = self.mc * self.intervals * self.cpi + 1
      self.ncon = self.intervals * self.n
    elif self.optimizer == TRAPEZOIDAL:
      self.nx_nodes = self.nu_nodes = self.intervals + 1
      self.ncon = self.intervals * self.n
    else:
      self.nx_nodes = self.nu_nodes = 2 * self.intervals + 1
      self.ncon = 2 * self.intervals * self.n
    self.nvars = self.nx_nodes * self.n + self.nu_nodes * self.m

  def desc(self, optimizer: Optional[int] = None, intervals: Optional[int] = None, cpi: Optional[int] = None,
           device="cuda") -> ML.MyrDesc:
    """device: where the NODE weights (if any) are placed -- "cuda" for the product path, "host" for the myr_host_* twins."""
    s = self.system
    extra = {}
    if getattr(s, "theta", None) is not None:  # NodeSystem
      if device == "host":
        th = np.ascontiguousarray(s.theta, dtype=np.float64)
        extra = dict(hidden=s.hidden, theta_ptr=th.ctypes.data, theta_doubles=th.size, keepalive=th)
      else:
        th = s.theta_device(torch.device(device))
        extra = dict(hidden=s.hidden, theta_ptr=th.data_ptr(), theta_doubles=th.numel(), keepalive=th)
    return ML.make_desc(s.device_name, self.optimizer if optimizer is None else optimizer, self.method,
                        self.intervals if intervals is None else intervals, self.cpi if cpi is None else cpi,
                        T=float(s.T), params=list(s.params), terminal_cost=bool(s.terminal_cost), **extra)

  def unravel(self, z):
    """ravel_pytree((x, u)) inverse; works for numpy arrays and torch tensors, batched or not."""
    nx = self.nx_nodes * self.n
    x = z[..., :nx].reshape(z.shape[:-1] + (self.nx_nodes, self.n))
    u = z[..., nx:].reshape(z.shape[:-1] + (self.nu_nodes, self.m))
    return x, u


def _linspace_rows(a: torch.Tensor, b: torch.Tensor, num: int) -> torch.Tensor:
  """jnp.linspace(a, b, num) per batch row: a + k * ((b - a) / (num - 1)), last entry exactly b.  a,b: [B] -> [B,num]"""
  k = torch.arange(num, dtype=torch.float64, device=a.device)
  step = (b - a) / (num - 1)
  out = a[:, None] + k[None, :] * step[:, None]
  out[:, -1] = b
  return out


def _rollout_guess(tr: Transcription, x0s: torch.Tensor, steps: int) -> torch.Tensor:
  """integrate_time_independent(dynamics, x_0, zeros, T/steps, steps, method)[1] -> [B, steps+1, n]"""
  eng = Engine(tr.desc(optimizer=TRAPEZOIDAL, intervals=steps, cpi=1, device=x0s.device))
  u = torch.zeros(x0s.shape[0], steps + 1, tr.m, dtype=torch.float64, device=x0s.device)
  xs, _ = eng.rollout_cost(u, x0s, want_states=True)
  return xs


def build_batch(tr: Transcription, x0s: torch.Tensor):
  """-> (z0, lb, ub), each [B, nvars] float64 on x0s.device.

  Systems whose target state is fully specified (the guess is a straight line from x0 to x_T, no rollout) take a
  vectorised path: the bounds of an instance differ from those of any other only in the start-state rows, so one template
  per (transcription, device) is built by the general code below and the start states are written into copies of it --
  a dozen launches instead of one per state component and bound row (it is inside the end-to-end timed region)."""
  xT = None if tr.system.x_T is None else list(tr.system.x_T)
  if xT is None or any(v is None for v in xT) or x0s.shape[0] == 0:
    return _build_batch_general(tr, x0s)
  cache = tr.__dict__.setdefault("_bb_cache", {})
  key = str(x0s.device)
  if key not in cache:
    _, lb1, ub1 = _build_batch_general(tr, x0s[:1])
    f64 = dict(dtype=torch.float64, device=x0s.device)
    cache[key] = (lb1[0].clone(), ub1[0].clone(), torch.arange(tr.nx_nodes, **f64),
                  torch.as_tensor(np.asarray(xT, dtype=np.float64), **f64))
  lb_t, ub_t, k, xt = cache[key]
  B, n, L = x0s.shape[0], tr.n, tr.nx_nodes
  step = (xt[None, :] - x0s) / (L - 1)                           # jnp.linspace(x0_i, xT_i, L): a + k * ((b - a) / (L - 1)),
  xg = x0s[:, None, :] + k[None, :, None] * step[:, None, :]     # last entry exactly b (same arithmetic as _linspace_rows)
  xg[:, -1, :] = xt
  z0 = torch.zeros(B, lb_t.shape[0], dtype=torch.float64, device=x0s.device)
  z0[:, :L * n] = xg.reshape(B, -1)
  lb = lb_t.unsqueeze(0).repeat(B, 1)
  ub = ub_t.unsqueeze(0).repeat(B, 1)
  lb[:, :n] = x0s
  ub[:, :n] = x0s
  return z0, lb, ub


def _build_batch_general(tr: Transcription, x0s: torch.Tensor):
  s = tr.system
  B = x0s.shape[0]
  dev = x0s.device
  n, m = tr.n, tr.m
  f64 = dict(dtype=torch.float64, device=dev)
  xT = None if s.x_T is None else list(s.x_T)
  bounds = torch.as_tensor(np.asarray(s.bounds, dtype=np.float64), **f64)
  L = tr.nx_nodes

  # ---- state guess
  if tr.optimizer == HERMITE_SIMPSON:
    if xT is not None:
      xg = torch.stack([_linspace_rows(x0s[:, i], torch.full((B,), float(xT[i]), **f64), L) for i in range(n)], dim=2)
    else:
      xg = torch.full((B, L, n), 0.1, **f64)
  else:
    steps = tr.intervals
    rolled = None
    if xT is None or any(v is None for v in xT):
      rolled = _rollout_guess(tr, x0s, steps)
    if xT is None:
      xg = rolled
    else:
      cols = []
      for i in range(n):
        if xT[i] is not None:
          cols.append(_linspace_rows(x0s[:, i], torch.full((B,), float(xT[i]), **f64), L))
        else:
          cols.append(rolled[:, :, i])
      xg = torch.stack(cols, dim=2)
  ug = torch.zeros(B, tr.nu_nodes, m, **f64)
  z0 = torch.cat([xg.reshape(B, -1), ug.reshape(B, -1)], dim=1).contiguous()

  # ---- bounds
  xb = bounds[:n].unsqueeze(0).unsqueeze(0).expand(B, L, n, 2).clone()
  xb[:, 0, :, 0] = x0s
  xb[:, 0, :, 1] = x0s
  if xT is not None:
    if tr.optimizer == TRAPEZOIDAL:
      # x_bounds[-control_shape] = x_T for every component (trapezoidal.py:70-71)
      xt = torch.as_tensor(np.asarray(xT, dtype=np.float64), **f64)
      xb[:, L - m, :, 0] = xt
      xb[:, L - m, :, 1] = xt
    else:
      for i in range(n):
        if xT[i] is not None:
          xb[:, -1, i, :] = float(xT[i])
  # control bounds are filled control-major; equal to time-major for m == 1
  Lu = tr.nu_nodes
  ub_rows = torch.empty(Lu * m, 2, **f64)
  for i in range(m, 0, -1):
    ub_rows[(m - i) * Lu:(m - i + 1) * Lu] = bounds[-i]
  lb = torch.cat([xb[..., 0].reshape(B, -1), ub_rows[:, 0].unsqueeze(0).expand(B, -1)], dim=1).contiguous()
  ub = torch.cat([xb[..., 1].reshape(B, -1), ub_rows[:, 1].unsqueeze(0).expand(B, -1)], dim=1).contiguous()
  return z0, lb, ub


def sample_x0(system: FiniteHorizonControlSystem, B: int, seed: int = 2019, spread: float = 0.1, device="cpu") -> torch.Tensor:
  """x0_b = clip(x_0 + spread * N(0, I), state bounds) -- the reference's start-state perturbation
  (hp.start_spread, myriad/config.py:80; recipe myriad/utils.py:412-419), drawn with a CPU torch generator
  so that row b is the same for every batch size / shard.  Row 0 is the unperturbed x_0."""
  g = torch.Generator(device="cpu").manual_seed(seed)
  n = system.state_size
  noise = torch.randn(B, n, generator=g, dtype=torch.float64)
  noise[0] = 0.0
  x0 = torch.as_tensor(np.asarray(system.x_0, dtype=np.float64)).unsqueeze(0) + spread * noise
  b = torch.as_tensor(np.asarray(system.bounds, dtype=np.float64))[:n]
  x0 = torch.minimum(torch.maximum(x0, b[:, 0]), b[:, 1])
  return x0.to(device)


def block_index_maps(tr: Transcription):
  """Index maps between the compact block layouts of the C ABI and the reference's flat layouts.

  Returns (zidx [nodes, nw], cidx [stages, nc], stage_node [stages, stage_nodes]) as numpy int64 arrays:
  zidx[q, i] = position in ravel_pytree((x, u)) of variable i of node q; cidx[j, r] = position in the
  reference constraint vector of row r of stage j; stage_node[j, k] = node whose block sits in slot k of
  stage j's row of Jblk."""
  n, m = tr.n, tr.m
  nw = n + m
  if tr.optimizer == TRAPEZOIDAL:
    Q, S, nc, sn = tr.intervals + 1, tr.intervals, n, 2
    cidx = np.arange(S * n).reshape(S, n)
    stage_node = np.stack([np.arange(S), np.arange(S) + 1], axis=1)
  elif tr.optimizer == HERMITE_SIMPSON:
    Q, S, nc, sn = 2 * tr.intervals + 1, tr.intervals, 2 * n, 3
    cidx = np.concatenate([np.arange(S * n).reshape(S, n), S * n + np.arange(S * n).reshape(S, n)], axis=1)
    stage_node = np.stack([2 * np.arange(S), 2 * np.arange(S) + 1, 2 * np.arange(S) + 2], axis=1)
  else:
    raise NotImplementedError("block maps for SHOOTING")
  q = np.arange(Q)[:, None]
  i = np.arange(nw)[None, :]
  zidx = np.where(i < n, q * n + i, Q * n + q * m + (i - n))
  return zidx.astype(np.int64), cidx.astype(np.int64), stage_node.astype(np.int64)


def dense_jacobian(tr: Transcription, Jblk: torch.Tensor) -> torch.Tensor:
  """Compact block Jacobian [B, stages, stage_nodes, nc, nw] -> dense [B, ncon, nvars] (what jax.jacrev
  returns in the reference, myriad/nlp_solvers/__init__.py:37)."""
  if tr.optimizer == SHOOTING:
    # Jblk [B, K, n+1, ncol]: rows 0..n-1 = d px_k / d (xs[k], controls of interval k); d c_k / d xs[k+1] = -I
    B, K, _, ncol = Jblk.shape
    n, m = tr.n, tr.m
    M = tr.mc * tr.cpi
    out = torch.zeros(B, tr.ncon, tr.nvars, dtype=Jblk.dtype, device=Jblk.device)
    ubase = (K + 1) * n
    for k in range(K):
      out[:, k * n:(k + 1) * n, k * n:(k + 1) * n] += Jblk[:, k, :n, :n]
      out[:, k * n:(k + 1) * n, ubase + k * M * m: ubase + (k * M + M + 1) * m] += Jblk[:, k, :n, n:]
      out[:, k * n:(k + 1) * n, (k + 1) * n:(k + 2) * n] -= torch.eye(n, dtype=Jblk.dtype, device=Jblk.device)
    return out
  zidx, cidx, stage_node = block_index_maps(tr)
  B = Jblk.shape[0]
  S, sn, nc, nw = Jblk.shape[1:]
  rows = torch.as_tensor(cidx, device=Jblk.device)[:, None, :, None].expand(S, sn, nc, nw)
  cols = torch.as_tensor(zidx[stage_node], device=Jblk.device)[:, :, None, :].expand(S, sn, nc, nw)
  flat = (rows * tr.nvars + cols).reshape(-1)
  out = torch.zeros(B, tr.ncon * tr.nvars, dtype=Jblk.dtype, device=Jblk.device)
  out.index_add_(1, flat, Jblk.reshape(B, -1))
  return out.reshape(B, tr.ncon, tr.nvars)
