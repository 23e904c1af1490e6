"""Thin torch-facing layer over the C ABI: torch CUDA fp64 tensors are the device buffers
(``tensor.data_ptr()``), the current torch CUDA stream is the launch stream.  No math happens here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib as ML


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
  for t in ts:
    if t is None:
      continue
    if not t.is_cuda:
      raise ML.MyriadError("myriad_b200 kernels take CUDA tensors; there is no CPU fallback")
    if t.dtype != torch.float64 or not t.is_contiguous():
      raise ML.MyriadError("expected contiguous float64 tensors")


@dataclass
class EvalResult:
  f: torch.Tensor
  grad: torch.Tensor
  c: torch.Tensor
  Jblk: torch.Tensor
  Hblk: Optional[torch.Tensor]


class Engine:
  """Holds a descriptor, its sizes and reusable device workspaces."""

  def __init__(self, desc: ML.MyrDesc):
    self.desc = desc
    self.sizes = ML.problem_sizes(desc)
    self._ws = None
    s = self.sizes
    if desc.optimizer == ML.OPT_SHOOTING:
      # per interval: (n + 1) x (n + (mc*cpi + 1) m); row n = gradient of the interval's integrated cost
      self.jblk_shape = (desc.intervals, s.n + 1, s.jac_block_doubles // (desc.intervals * (s.n + 1)))
    else:
      self.jblk_shape = (s.stages, s.stage_nodes, s.nc, s.nw)

  # ---- K1
  def eval(self, z: torch.Tensor, lam: Optional[torch.Tensor] = None, hessian: bool = False, out: Optional[EvalResult] = None,
           jac: bool = True) -> EvalResult:
    """jac=False: objective and constraints only (grad / Jblk are None).  The stand-alone K1 stages an instance's whole block
    Jacobian in shared memory; the one configuration where that does not fit (NODE dynamics + Hermite-Simpson at N ~ 100:
    myr_eval returns MYR_E_UNSUPPORTED -> NotImplementedError) still evaluates f and c this way."""
    _need_cuda(z, lam)
    s = self.sizes
    B = z.shape[0]
    assert z.shape[1] == s.nvars
    dev = z.device
    if out is None:
      out = EvalResult(
        f=torch.empty(B, dtype=torch.float64, device=dev),
        grad=torch.empty(B, s.nvars, dtype=torch.float64, device=dev) if jac else None,
        c=torch.empty(B, s.ncon, dtype=torch.float64, device=dev),
        Jblk=torch.empty((B,) + self.jblk_shape, dtype=torch.float64, device=dev) if jac else None,
        Hblk=torch.empty(B, s.nodes, s.nw * (s.nw + 1) // 2, dtype=torch.float64, device=dev) if hessian else None)
    if hessian and lam is None:
      lam = torch.zeros(B, s.ncon, dtype=torch.float64, device=dev)
    ML.check(ML.lib().myr_eval(C.byref(self.desc), B, _ptr(z), _ptr(lam), _ptr(out.f), _ptr(out.grad) if jac else None, _ptr(out.c),
                               _ptr(out.Jblk) if jac else None, _ptr(out.Hblk) if hessian else None, _stream()))
    return out

  def jtvec(self, Jblk: torch.Tensor, lam: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """J^T lam ([B, nvars], reference layout) from the compact block Jacobian of ``eval`` (myr_jtvec)"""
    _need_cuda(Jblk, lam)
    B = lam.shape[0]
    if out is None:
      out = torch.empty(B, self.sizes.nvars, dtype=torch.float64, device=lam.device)
    ML.check(ML.lib().myr_jtvec(C.byref(self.desc), B, _ptr(Jblk), _ptr(lam), _ptr(out), _stream()))
    return out

  def workspace(self, B: int, device) -> torch.Tensor:
    need = ML.workspace_doubles(self.sizes, B)
    if self._ws is None or self._ws.numel() < need or self._ws.device != device:
      self._ws = torch.empty(need, dtype=torch.float64, device=device)
    return self._ws

  # ---- K2
  def kkt_solve(self, Hblk, Jblk, sigma, rhs_z, rhs_c, delta_w=0.0, delta_c=0.0):
    _need_cuda(Hblk, Jblk, sigma, rhs_z, rhs_c)
    s = self.sizes
    B = sigma.shape[0]
    dev = sigma.device
    dz = torch.empty(B, s.nvars, dtype=torch.float64, device=dev)
    dlam = torch.empty(B, s.ncon, dtype=torch.float64, device=dev)
    ok = torch.empty(B, dtype=torch.int32, device=dev)
    ws = self.workspace(B, dev)
    ML.check(ML.lib().myr_kkt_solve(C.byref(self.desc), B, _ptr(Hblk), _ptr(Jblk), _ptr(sigma), _ptr(rhs_z), _ptr(rhs_c),
                                    float(delta_w), float(delta_c), _ptr(dz), _ptr(dlam), _ptr(ok), _ptr(ws), ws.numel(), _stream()))
    return dz, dlam, ok

  # ---- K3
  def ipm_solve(self, z0, lb, ub, max_iter=1000, tol=1e-8, acceptable_tol=1e-6, mu_init=0.1, out=None):
    _need_cuda(z0, lb, ub)
    s = self.sizes
    B = z0.shape[0]
    dev = z0.device
    if out is None:
      out = {
        "z": torch.empty(B, s.nvars, dtype=torch.float64, device=dev),
        "lam": torch.empty(B, s.ncon, dtype=torch.float64, device=dev),
        "zL": torch.empty(B, s.nvars, dtype=torch.float64, device=dev),
        "zU": torch.empty(B, s.nvars, dtype=torch.float64, device=dev),
        "obj": torch.empty(B, dtype=torch.float64, device=dev),
        "kkt_err": torch.empty(B, dtype=torch.float64, device=dev),
        "con_inf": torch.empty(B, dtype=torch.float64, device=dev),
        "status": torch.empty(B, dtype=torch.int32, device=dev),
        "iters": torch.empty(B, dtype=torch.int32, device=dev),
      }
    o = ML.MyrIpmOpts()
    o.max_iter = int(max_iter)
    o.tol = float(tol)
    o.acceptable_tol = float(acceptable_tol)
    o.mu_init = float(mu_init)
    ws = self.workspace(B, dev)
    ML.check(ML.lib().myr_ipm_solve(C.byref(self.desc), C.byref(o), B, _ptr(z0), _ptr(lb), _ptr(ub), _ptr(out["z"]), _ptr(out["lam"]),
                                    _ptr(out["zL"]), _ptr(out["zU"]), _ptr(out["obj"]), _ptr(out["kkt_err"]), _ptr(out["con_inf"]),
                                    _ptr(out["status"]), _ptr(out["iters"]), _ptr(ws), ws.numel(), _stream()))
    return out

  # ---- verification rollout
  def rollout_cost(self, u: torch.Tensor, x0: torch.Tensor, want_states: bool = True):
    _need_cuda(u, x0)
    B, rows, m = u.shape
    n = x0.shape[1]
    steps = self.desc.intervals * (self.desc.controls_per_interval if self.desc.optimizer == ML.OPT_SHOOTING else 1)
    xs = torch.empty(B, steps + 1, n, dtype=torch.float64, device=u.device) if want_states else None
    cost = torch.empty(B, dtype=torch.float64, device=u.device)
    ML.check(ML.lib().myr_rollout_cost(C.byref(self.desc), B, rows, _ptr(u), _ptr(x0), _ptr(xs), _ptr(cost), _stream()))
    return xs, cost
