"""FBSM: mirror of myriad/trajectory_optimizers/forward_backward_sweep.py:20-158 (and IndirectMethodOptimizer,
myriad/trajectory_optimizers/base.py:106-141) on the batched sweep kernel (csrc/fbsm.cuh, myr_fbsm_solve).

``solve()`` keeps the reference's result -- {'x', 'u', 'adj'} as (N+1, n|m) arrays for ``system.x_0``; ``solve_batch(x0)``
is the capability the reference lacks: one launch for many start states.  No math happens here: the sweeps, the control
update, the stopping rule and the secant iteration for a terminal state value all run in the kernel.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from myriad_b200 import _lib as ML
from myriad_b200.config import Config, HParams

# What the reference's IndirectFHCS subclasses carry besides the data of FiniteHorizonControlSystem
# (myriad/systems/lenhart/*.py): the terminal adjoint adj_T and WHICH bounds rows optim_characterization clamps with.
#   adj_T: None | list | callable(system) -> list
#   rows:  indices into system.bounds (SIMPLECASE really uses row 0, the -inf..inf state row: simple_case.py:63)
INDIRECT = {
  "SIMPLECASE": dict(adj_T=None, rows=[0]),
  "SIMPLECASEWITHBOUNDS": dict(adj_T=None, rows=[-1]),
  "CANCERTREATMENT": dict(adj_T=None, rows=[-1]),
  "MOULDFUNGICIDE": dict(adj_T=None, rows=[-1]),
  "BIOREACTOR": dict(adj_T=None, rows=[-1]),
  "GLUCOSE": dict(adj_T=None, rows=[-1]),  # not clamped at all (glucose.py:101-105); the row is ignored by the kernel
  "HARVEST": dict(adj_T=None, rows=[-1]),
  "TIMBERHARVEST": dict(adj_T=None, rows=[-1]),
  "EPIDEMICSEIRN": dict(adj_T=None, rows=[-1]),
  "HIVTREATMENT": dict(adj_T=None, rows=[-1]),
  "BACTERIA": dict(adj_T=lambda s: [s.params[3]], rows=[-1]),  # adj_T = [C] (bacteria.py:49)
  "PREDATORPREY": dict(adj_T=[1.0, 0.0, 0.0], rows=[-1]),      # predator_prey.py:68
  "BEARPOPULATIONS": dict(adj_T=None, rows=[-2, -1]),
  "INVASIVEPLANT": dict(adj_T=[1.0] * 5, rows=[-1] * 5),  # discrete; every control clamps with the last row (invasive_plant.py:90)
}


class IndirectMethodOptimizer(object):
  """myriad/trajectory_optimizers/base.py:106-141"""
  require_adj: bool = True

  def solve(self):
    raise NotImplementedError

  def stopping_criterion(self, x_iter, u_iter, adj_iter, delta: float = 0.001) -> bool:
    """The rule the kernel applies after every sweep (base.py:128-141); kept for user code."""
    x, old_x = x_iter
    u, old_u = u_iter
    adj, old_adj = adj_iter
    stop_x = np.abs(x).sum(axis=0) * delta - np.abs(x - old_x).sum(axis=0)
    stop_u = np.abs(u).sum(axis=0) * delta - np.abs(u - old_u).sum(axis=0)
    stop_adj = np.abs(adj).sum(axis=0) * delta - np.abs(adj - old_adj).sum(axis=0)
    return bool(np.min(np.hstack((stop_u, stop_x, stop_adj))) < 0)


class FBSM(IndirectMethodOptimizer):
  def __init__(self, hp: HParams, cfg: Config, system) -> None:
    self.hp, self.cfg, self.system = hp, cfg, system
    name = getattr(system, "device_name", "")
    if name not in INDIRECT:
      raise NotImplementedError(f"system {name or type(system).__name__} has no adj_ODE / optim_characterization: "
                                "FBSM needs an indirect (Lenhart) system")
    self.N = int(hp.fbsm_intervals)
    self.h = float(system.T) / self.N
    self.discrete = bool(getattr(system, "discrete", False))
    if self.discrete:  # forward_backward_sweep.py:33-35
      self.N, self.h = int(system.T), 1
    self.u_rows = self.N if self.discrete else self.N + 1
    n, m = system.state_size, system.control_size
    info = INDIRECT[name]
    adj_T = info["adj_T"](system) if callable(info["adj_T"]) else info["adj_T"]
    self.adj_T = None if adj_T is None else np.asarray(adj_T, dtype=np.float64)
    b = np.asarray(system.bounds, dtype=np.float64)
    self.char_lb = np.ascontiguousarray(b[info["rows"], 0])
    self.char_ub = np.ascontiguousarray(b[info["rows"], 1])
    # the reference's guesses (forward_backward_sweep.py:40-50), kept as attributes like there
    self.x_guess = np.vstack((np.asarray(system.x_0, dtype=np.float64), np.zeros((self.N, n))))
    self.u_guess = np.zeros((self.u_rows, m))
    self.adj_guess = np.zeros((self.N + 1, n)) if self.adj_T is None else np.vstack((np.zeros((self.N, n)), self.adj_T))
    self.t_interval = np.linspace(0, system.T, num=self.N + 1).reshape(-1, 1)
    self.guess = np.concatenate([self.x_guess.ravel(), self.u_guess.ravel(), self.adj_guess.ravel()])
    self.x_bounds, self.u_bounds = b[:-1], b[-1:]
    self.bounds = np.vstack((self.x_bounds, self.u_bounds))
    # one state with a terminal value => secant iteration on its terminal adjoint (forward_backward_sweep.py:57-70)
    self.terminal_cdtion, self.term_cdtion_state, self.term_value = False, -1, 0.0
    if system.x_T is not None:
      count = 0
      for idx, v in enumerate(system.x_T):
        if v is not None:
          self.terminal_cdtion, self.term_cdtion_state, self.term_value = True, idx, float(v)
          count += 1
        if count > 1:
          raise NotImplementedError("Multiple states with terminal condition not supported yet")
    self.opts = ML.MyrFbsmOpts(max_iter=0, max_secant=0, term_state=self.term_cdtion_state, reserved=0, delta=0.0,
                               secant_tol=0.0, term_value=self.term_value,
                               guess_a=float(getattr(system, "guess_a", 0.0)), guess_b=float(getattr(system, "guess_b", 0.0)))
    self.desc = ML.make_desc(name, ML.OPT_SHOOTING, "RK4", self.N, 1, T=float(system.T), params=list(system.params),
                             terminal_cost=bool(system.terminal_cost))

  def _call(self, fn, x0, x, u, adj, iters, status, *stream):
    p = lambda a: C.c_void_p(a.data_ptr())
    hp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    ML.check(fn(C.byref(self.desc), C.byref(self.opts), int(x0.shape[0]), p(x0), hp(self.adj_T), hp(self.char_lb),
                hp(self.char_ub), p(x), p(u), p(adj), p(iters), p(status), *stream))

  def solve_batch(self, x0, device=None) -> Dict[str, torch.Tensor]:
    """x0: (B, n) start states (host or device).  Returns device tensors x, adj: (B, N+1, n), u: (B, N+1, m) (views of the
    kernel's time-major storage; N rows of u for a discrete system), iters, status: (B,)."""
    if not torch.cuda.is_available():
      raise ML.MyriadError("FBSM.solve_batch runs the CUDA sweep kernel: a CUDA device is required (no CPU fallback)")
    dev = torch.device(device or "cuda")
    n, m = self.system.state_size, self.system.control_size
    x0 = torch.as_tensor(np.asarray(x0, dtype=np.float64) if not torch.is_tensor(x0) else x0, dtype=torch.float64)
    x0 = x0.reshape(-1, n).to(dev, non_blocking=True).contiguous()
    B = x0.shape[0]
    x = torch.empty(self.N + 1, n, B, dtype=torch.float64, device=dev)
    u = torch.empty(self.u_rows, m, B, dtype=torch.float64, device=dev)
    adj = torch.empty(self.N + 1, n, B, dtype=torch.float64, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
      self._call(ML.lib().myr_fbsm_solve, x0, x, u, adj, iters, status, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return {"x": x.permute(2, 0, 1), "u": u.permute(2, 0, 1), "adj": adj.permute(2, 0, 1), "iters": iters, "status": status}

  def solve(self) -> Dict[str, np.ndarray]:
    """forward_backward_sweep.py:91-116: {'x', 'u', 'adj'} for system.x_0"""
    r = self.solve_batch(np.asarray(self.system.x_0, dtype=np.float64).reshape(1, -1))
    out = {k: r[k][0].cpu().numpy() for k in ("x", "u", "adj")}
    self.x_guess, self.u_guess, self.adj_guess = out["x"], out["u"], out["adj"]
    return out

  def host_solve_batch(self, x0) -> Dict[str, np.ndarray]:
    """The same templates compiled for the host (myr_host_fbsm_solve): debugging / CI without a GPU only."""
    n, m = self.system.state_size, self.system.control_size
    x0 = torch.as_tensor(np.asarray(x0, dtype=np.float64)).reshape(-1, n).contiguous()
    B = x0.shape[0]
    x = torch.empty(self.N + 1, n, B, dtype=torch.float64)
    u = torch.empty(self.u_rows, m, B, dtype=torch.float64)
    adj = torch.empty(self.N + 1, n, B, dtype=torch.float64)
    iters = torch.empty(B, dtype=torch.int32)
    status = torch.empty(B, dtype=torch.int32)
    self._call(ML.lib().myr_host_fbsm_solve, x0, x, u, adj, iters, status)
    return {"x": x.permute(2, 0, 1).numpy(), "u": u.permute(2, 0, 1).numpy(), "adj": adj.permute(2, 0, 1).numpy(),
            "iters": iters.numpy(), "status": status.numpy()}

  def solve_batch_sharded(self, x0_all, host: bool = False) -> Dict[str, torch.Tensor]:
    """One process per GPU: this rank solves its contiguous share of the start states (myriad_b200.distributed.shard_range)
    and ONE all_gather hands every rank all trajectories in global row order.  Instances are independent, so there is no
    data-path collective.  ``host=True`` uses the host build (CPU CI with the gloo backend)."""
    from myriad_b200 import distributed as D
    rank, world, _ = D.world()
    n, m = self.system.state_size, self.system.control_size
    x0_all = torch.as_tensor(np.asarray(x0_all, dtype=np.float64) if not torch.is_tensor(x0_all) else x0_all).reshape(-1, n)
    lo, hi = D.shard_range(x0_all.shape[0], rank, world)
    r = self.host_solve_batch(x0_all[lo:hi].cpu().numpy()) if host else self.solve_batch(x0_all[lo:hi])
    t = lambda a: torch.as_tensor(a)
    B = hi - lo
    packed = torch.cat([t(r["x"]).reshape(B, -1), t(r["u"]).reshape(B, -1), t(r["adj"]).reshape(B, -1),
                        t(r["iters"]).double()[:, None], t(r["status"]).double()[:, None]], dim=1).contiguous()
    allp = D.gather_solutions(packed)
    nx, nu = (self.N + 1) * n, self.u_rows * m
    Bt = allp.shape[0]
    return {"x": allp[:, :nx].reshape(Bt, self.N + 1, n), "u": allp[:, nx:nx + nu].reshape(Bt, self.u_rows, m),
            "adj": allp[:, nx + nu:nx + nu + nx].reshape(Bt, self.N + 1, n), "iters": allp[:, -2].to(torch.int32),
            "status": allp[:, -1].to(torch.int32)}
