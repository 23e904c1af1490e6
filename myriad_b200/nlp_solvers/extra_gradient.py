"""EXTRAGRADIENT solver: mirror of myriad/nlp_solvers/extra_gradient.py:10-82, batched over problem instances.

The reference iterates on the Lagrangian L(x, lam) = f(x) + lam . c(x) with bounds handled by clipping (:27-33):

    x_bar   = clip(x - eta_x * grad_x L(x,     lam))
    x_new   = clip(x - eta_x * grad_x L(x_bar, lam))
    lam_new = lam + eta_v * c(x_new)

Here grad_x L = grad f + J^T lam comes from the K1 kernel (myr_eval: objective gradient, constraints and the compact
block Jacobian in one launch) and the VJP kernel (myr_jtvec); the three cheap vector updates are torch elementwise ops
on the same device buffers.  Every ``check_every`` (1000) iterations the step sizes decay by 0.999 (:54-56) and an
instance whose last step moved no variable by more than ``atol`` is declared converged (:50-52) and frozen; the loop
ends when all instances are, or after ``max_iter`` steps.  lam starts at ones (:76).
"""
from __future__ import annotations

from typing import Dict

import torch

# myriad/defaults.py:5-18
LEARNING_RATES = {
  "PENDULUM": {"eta_x": 1e-1, "eta_v": 1e-3},
  "CANCERTREATMENT": {"eta_x": 1e-1, "eta_v": 1e-3},
  "CARTPOLE": {"eta_x": 1e-2, "eta_v": 1e-4},
}


def extra_gradient_batch(hp, eng, z0: torch.Tensor, lb: torch.Tensor, ub: torch.Tensor, max_iter: int = 30_000,
                         eta_x: float = None, eta_v: float = None, atol: float = 1e-6, check_every: int = 1000) -> Dict[str, torch.Tensor]:
  rates = LEARNING_RATES.get(hp.system.name, {}) if hp is not None else {}
  eta_x = rates.get("eta_x", 1e-1) if eta_x is None else eta_x   # extra_gradient.py:17-18
  eta_v = rates.get("eta_v", 1e-3) if eta_v is None else eta_v
  B = z0.shape[0]
  s = eng.sizes
  x = z0.clone()
  lam = torch.ones(B, s.ncon, dtype=torch.float64, device=z0.device)
  x_old = x + 20.0
  done = torch.zeros(B, dtype=torch.bool, device=z0.device)
  r = None
  jt = torch.empty_like(x)
  it_done = torch.zeros(B, dtype=torch.int32, device=z0.device)
  for i in range(int(max_iter)):
    if i % check_every == 0:
      done = done | ((x_old - x).abs().amax(dim=1) <= atol)
      if bool(done.all()):
        break
      eta_x *= 0.999
      eta_v *= 0.999
    x_old = x
    r = eng.eval(x, out=r)
    eng.jtvec(r.Jblk, lam, out=jt)
    x_bar = torch.minimum(torch.maximum(x - eta_x * (r.grad + jt), lb), ub)
    r = eng.eval(x_bar, out=r)
    eng.jtvec(r.Jblk, lam, out=jt)
    x_new = torch.minimum(torch.maximum(x - eta_x * (r.grad + jt), lb), ub)
    r = eng.eval(x_new, out=r)
    lam_new = lam + eta_v * r.c
    keep = done[:, None]
    x = torch.where(keep, x, x_new)
    lam = torch.where(keep, lam, lam_new)
    it_done += (~done).to(torch.int32)
  r = eng.eval(x, out=r)
  return {"z": x, "lam": lam, "obj": r.f.clone(), "status": torch.where(done, 0, -1).to(torch.int32), "iters": it_done,
          "kkt_err": torch.full((B,), float("nan"), dtype=torch.float64, device=z0.device),
          "con_inf": r.c.abs().amax(dim=1),
          "zL": torch.zeros_like(x), "zU": torch.zeros_like(x)}
