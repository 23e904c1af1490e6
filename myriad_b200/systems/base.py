"""System descriptors: the B200 engine's mirror of myriad/systems/base.py:11-111.

A system here is DATA (x_0, x_T, T, bounds, parameters) plus the name of its generated device code
(csrc/systems_gen.cuh, produced by tools/gen_systems.py from the same formulas).  The solver never traces Python
callables the way jax does: the kernels evaluate the generated code.  ``dynamics`` / ``cost`` below keep the reference's
call signatures (myriad/systems/base.py:45-73) for user code: they evaluate the SAME generated device code through the
C ABI (myr_dynamics), one point or a batch of points per call, on the current CUDA device.
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

  # ---- the reference's callables (base.py:45-73), evaluated by the generated device code
  def _desc(self, device="cuda"):
    from myriad_b200 import problems as PR
    return PR.Transcription(self, PR.TRAPEZOIDAL, "HEUN", 1, 1).desc(device=device)

  def _points(self, x_t, u_t, t):
    import torch
    x = torch.as_tensor(np.asarray(x_t, dtype=np.float64)).reshape(-1, self.state_size)
    B = x.shape[0]
    u = torch.as_tensor(np.asarray(u_t, dtype=np.float64)).reshape(-1, self.control_size)
    if u.shape[0] != B:
      u = u.expand(B, self.control_size)
    tt = None
    if t is not None:
      tt = torch.as_tensor(np.asarray(t, dtype=np.float64)).reshape(-1)
      if tt.shape[0] != B:
        tt = tt.expand(B)
    return x, u, tt, B

  def _dynamics_call(self, x_t, u_t, t, want_f: bool, want_g: bool):
    import ctypes as C
    import torch
    from myriad_b200 import _lib as ML
    if not torch.cuda.is_available():
      raise ML.MyriadError("system.dynamics / system.cost evaluate the generated device code: a CUDA device is required")
    x, u, tt, B = self._points(x_t, u_t, t)
    dev = torch.device("cuda")
    x, u = x.to(dev).contiguous(), u.to(dev).contiguous()
    tt = None if tt is None else tt.to(dev).contiguous()
    f = torch.empty(B, self.state_size, dtype=torch.float64, device=dev) if want_f else None
    g = torch.empty(B, dtype=torch.float64, device=dev) if want_g else None
    ptr = lambda a: None if a is None else C.c_void_p(a.data_ptr())
    d = self._desc()
    ML.check(ML.lib().myr_dynamics(C.byref(d), B, ptr(x), ptr(u), ptr(tt), ptr(f), ptr(g),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return f, g

  def dynamics(self, x_t, u_t, t=None):
    """x'(t) = f(x, u): one point (returns shape (n,)) or a batch of points ([B, n])."""
    f, _ = self._dynamics_call(x_t, u_t, t, True, False)
    out = f.cpu().numpy()
    return out[0] if np.ndim(x_t) <= 1 else out

  def cost(self, x_t, u_t, t=None):
    """running cost g(x, u, t): scalar for one point, [B] for a batch."""
    _, g = self._dynamics_call(x_t, u_t, t, False, True)
    out = g.cpu().numpy()
    return float(out[0]) if np.ndim(x_t) <= 1 else out

  def parametrized_dynamics(self, params, x_t, u_t, t=None):
    """dynamics of the same system class built with ``params`` (constructor keywords), e.g. cartpole.py:89-104"""
    return type(self)(**params).dynamics(x_t, u_t, t)

  def parametrized_cost(self, params, x_t, u_t, t=None):
    return type(self)(**params).cost(x_t, u_t, t)
