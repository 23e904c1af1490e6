"""Development model (NumPy, dense linear algebra) of the batched primal-dual interior-point
method that myriad_b200/csrc implements in CUDA.  NOT part of the product and NOT the oracle:
it is the executable specification the CUDA IPM (csrc/ipm.cuh) was written from, kept for
algorithm experiments and as a host twin in tests (iteration counts / final KKT error).

Algorithm = line-search primal-dual IPM (Waechter & Biegler 2006 structure, l1 merit function
instead of the filter, per-block convexification instead of inertia detection):

  min f(z)  s.t. c(z) = 0,  lb <= z <= ub   (lb == ub entries are eliminated)
"""
from __future__ import annotations

import dataclasses
import sys
import time

import numpy as np


@dataclasses.dataclass
class IpmOptions:
  max_iter: int = 1000
  tol: float = 1e-8
  mu_init: float = 0.1
  mu_min: float = 1e-11  # tol / 10 is applied on top
  kappa_eps: float = 10.0
  kappa_mu: float = 0.2
  theta_mu: float = 1.5
  tau_min: float = 0.99
  bound_push: float = 1e-2
  bound_frac: float = 1e-2
  bound_relax: float = 1e-8
  kappa_sigma: float = 1e10
  s_max: float = 100.0
  delta_min: float = 1e-20
  delta_0: float = 1e-4
  delta_max: float = 1e40
  delta_c: float = 0.0
  eta: float = 1e-4  # Armijo
  rho: float = 0.1  # penalty-parameter margin
  max_ls: int = 40
  hessian: str = "exact"  # exact | bfgs
  verbose: bool = False
  capture: list = None  # if a list: (W, J, Sig, rb, c, free, delta) per iteration is appended


def solve_ipm(fun, grad, con, jac, hess, z0, lb, ub, opt: IpmOptions = IpmOptions(), blocks=None):
  """hess(z, lam) -> dense Hessian of f + lam.c (None when opt.hessian == 'bfgs')."""
  nv = z0.shape[0]
  fixed = lb == ub
  free = ~fixed
  hasL = np.isfinite(lb) & free
  hasU = np.isfinite(ub) & free
  lbr = np.where(hasL, lb - opt.bound_relax * np.maximum(1, np.abs(lb)), lb)
  ubr = np.where(hasU, ub + opt.bound_relax * np.maximum(1, np.abs(ub)), ub)

  z = z0.copy()
  z[fixed] = lb[fixed]
  both = hasL & hasU
  pL = np.where(both, np.minimum(opt.bound_push * np.maximum(1, np.abs(lbr)), opt.bound_frac * (ubr - lbr)),
                opt.bound_push * np.maximum(1, np.abs(lbr)))
  pU = np.where(both, np.minimum(opt.bound_push * np.maximum(1, np.abs(ubr)), opt.bound_frac * (ubr - lbr)),
                opt.bound_push * np.maximum(1, np.abs(ubr)))
  z = np.where(hasL, np.maximum(z, lbr + pL), z)
  z = np.where(hasU, np.minimum(z, ubr - pU), z)

  zL = np.where(hasL, 1.0, 0.0)
  zU = np.where(hasU, 1.0, 0.0)
  c = con(z)
  nc = c.shape[0]
  lam = np.zeros(nc)
  mu = opt.mu_init
  nu = 1.0  # merit penalty
  Bk = None
  delta_last = 0.0
  it = 0
  status = -1
  hist = []

  def slacks(zz):
    sL = np.where(hasL, zz - lbr, 1.0)
    sU = np.where(hasU, ubr - zz, 1.0)
    return sL, sU

  def barrier(zz, mu_):
    sL, sU = slacks(zz)
    return fun(zz) - mu_ * (np.log(sL[hasL]).sum() + np.log(sU[hasU]).sum())

  f = fun(z); g = grad(z); J = jac(z)
  g_old_lag = None
  z_old = None
  while True:
    sL, sU = slacks(z)
    # -------- optimality error (IPOPT eq. 5/6 scaling)
    rd = g + J.T @ lam - zL + zU
    rd[fixed] = 0.0
    sd = max(opt.s_max, (np.abs(lam).sum() + zL.sum() + zU.sum()) / max(1, nc + hasL.sum() + hasU.sum())) / opt.s_max
    sc = max(opt.s_max, (zL.sum() + zU.sum()) / max(1, hasL.sum() + hasU.sum())) / opt.s_max

    def err(mu_):
      cL = np.abs(sL * zL - mu_)[hasL].max(initial=0.0)
      cU = np.abs(sU * zU - mu_)[hasU].max(initial=0.0)
      return max(np.abs(rd).max() / sd, np.abs(c).max(initial=0.0), max(cL, cU) / sc)

    e0 = err(0.0)
    hist.append((it, f, np.abs(c).max(initial=0), np.abs(rd).max(), mu, e0))
    if opt.verbose:
      print(f"it {it:3d} f={f:.10f} |c|={np.abs(c).max(initial=0):.2e} |rd|={np.abs(rd).max():.2e} mu={mu:.1e} E0={e0:.2e} nu={nu:.1e}")
    if e0 <= opt.tol:
      status = 0
      break
    if it >= opt.max_iter:
      status = -1
      break
    mu_floor = max(opt.mu_min, opt.tol / 10)
    while err(mu) <= opt.kappa_eps * mu and mu > mu_floor:
      mu = max(mu_floor, min(opt.kappa_mu * mu, mu ** opt.theta_mu))
    tau = max(opt.tau_min, 1 - mu)

    # -------- Hessian
    if opt.hessian == "exact":
      W = hess(z, lam)
    else:
      glag = g + J.T @ lam
      blks = blocks if blocks is not None else [np.arange(nv)]
      if Bk is None:
        Bk = np.eye(nv)
        first = [True] * len(blks)
      elif z_old is not None:
        s_all = z - z_old
        y_all = glag - (g_old + J_old.T @ lam)
        for bi, b in enumerate(blks):
          s = s_all[b]; y = y_all[b]
          if np.abs(s).max() < 1e-14:
            continue
          Bb = Bk[np.ix_(b, b)]
          sy = s @ y
          if first[bi] and sy > 1e-12:
            Bb = np.eye(len(b)) * min(max((y @ y) / sy, 1e-3), 1e6)
            first[bi] = False
          Bs = Bb @ s
          sBs = s @ Bs
          if sBs > 1e-300:
            theta = 1.0 if sy >= 0.2 * sBs else 0.8 * sBs / (sBs - sy)
            r = theta * y + (1 - theta) * Bs
            Bb = Bb - np.outer(Bs, Bs) / sBs + np.outer(r, r) / (s @ r)
          Bk[np.ix_(b, b)] = Bb
      W = Bk
    Sig = np.where(hasL, zL / sL, 0.0) + np.where(hasU, zU / sU, 0.0)
    # -------- KKT solve with convexification
    rb = g - np.where(hasL, mu / sL, 0.0) + np.where(hasU, mu / sU, 0.0) + J.T @ lam  # grad of barrier Lagrangian
    idx = np.where(free)[0]
    Jf = J[:, idx]
    delta = 0.0
    nf = len(idx)
    while True:
      H = W[np.ix_(idx, idx)] + np.diag(Sig[idx] + delta)
      K = np.block([[H, Jf.T], [Jf, -opt.delta_c * np.eye(nc)]])
      ev = np.linalg.eigvalsh(K)
      nneg = int((ev < 0).sum()); nzero = int((np.abs(ev) < 1e-12).sum())
      if nneg == nc and nzero == 0:
        break
      if delta == 0.0:
        delta = opt.delta_0 if delta_last == 0.0 else max(opt.delta_min, delta_last / 3.0)
        first_try = True
      else:
        delta = delta * (100.0 if delta_last == 0.0 else 8.0)
      if delta > opt.delta_max:
        raise RuntimeError("inertia correction failed")
    if delta > 0:
      delta_last = delta
    if opt.capture is not None:
      opt.capture.append(dict(W=W.copy(), J=J.copy(), Sig=Sig.copy(), rb=rb.copy(), c=c.copy(), free=free.copy(), delta=delta))
    sol = np.linalg.solve(K, -np.concatenate([rb[idx], c]))
    dzf = sol[:nf]; dlam = sol[nf:]
    dz = np.zeros(nv); dz[idx] = dzf
    dzL = np.where(hasL, mu / sL - zL - zL / sL * dz, 0.0)
    dzU = np.where(hasU, mu / sU - zU + zU / sU * dz, 0.0)
    # -------- fraction to the boundary
    def amax(v, dv, mask):
      m = mask & (dv < 0)
      return min(1.0, (-tau * v[m] / dv[m]).min(initial=1.0))
    a_pr = min(amax(sL, dz, hasL), amax(sU, -dz, hasU))
    a_du = min(amax(zL, dzL, hasL), amax(zU, dzU, hasU))
    # -------- l1 merit line search
    gb = g - np.where(hasL, mu / sL, 0.0) + np.where(hasU, mu / sU, 0.0)
    dphi = gb @ dz
    c1 = np.abs(c).sum()
    dHd = dzf @ (H @ dzf)
    if c1 > 0:
      nu_trial = (dphi + 0.5 * max(dHd, 0.0)) / ((1 - opt.rho) * c1)
      if nu < nu_trial:
        nu = nu_trial + 1.0
    D = dphi - nu * c1
    phi0 = barrier(z, mu) + nu * c1
    a = a_pr
    ok = False
    for ls in range(opt.max_ls):
      zt = z + a * dz
      ct = con(zt)
      phit = barrier(zt, mu) + nu * np.abs(ct).sum()
      if np.isfinite(phit) and phit <= phi0 + opt.eta * a * D:
        ok = True
        break
      a *= 0.5
    if not ok:
      status = -2
      break
    z_old, g_old, J_old = z, g, J
    z = zt
    lam = lam + a * dlam
    zL = zL + a_du * dzL
    zU = zU + a_du * dzU
    sL, sU = slacks(z)
    zL = np.where(hasL, np.clip(zL, mu / (opt.kappa_sigma * sL), opt.kappa_sigma * mu / sL), 0.0)
    zU = np.where(hasU, np.clip(zU, mu / (opt.kappa_sigma * sU), opt.kappa_sigma * mu / sU), 0.0)
    c = ct
    f = fun(z); g = grad(z); J = jac(z)
    it += 1
    if opt.verbose:
      print(f"      alpha={a:.3e} a_du={a_du:.3e} delta={delta:.1e} ls={ls}")
  return dict(z=z, lam=lam, zL=zL, zU=zU, f=f, status=status, iters=it, hist=hist)


def node_blocks(tr):
  """Index sets of the separable pieces of the Lagrangian: one per collocation node; the whole
  vector for shooting."""
  from oracle.transcription import Shooting
  if isinstance(tr, Shooting):
    return None
  n, m, L = tr.n, tr.m, tr.nx_nodes
  return [np.concatenate([np.arange(k * n, (k + 1) * n), L * n + np.arange(k * m, (k + 1) * m)]) for k in range(L)]


def solve_transcription(tr, opt: IpmOptions = IpmOptions(), guess=None):
  from oracle import nlp
  z0 = tr.guess if guess is None else guess
  return solve_ipm(lambda z: float(tr.objective(z)), lambda z: nlp.objective_grad(tr, z),
                   lambda z: tr.constraints(z), lambda z: nlp.constraints_jac(tr, z),
                   (lambda z, lam: nlp.lagrangian_hessian(tr, z, lam)) if opt.hessian == "exact" else None,
                   z0, tr.bounds[:, 0].copy(), tr.bounds[:, 1].copy(), opt, blocks=node_blocks(tr))


if __name__ == "__main__":
  sys.path.insert(0, ".")
  from oracle.systems import make_system
  from oracle.transcription import make_transcription
  name, optz, intervals, cpi, meth, quad = sys.argv[1:7]
  hessian = sys.argv[7] if len(sys.argv) > 7 else "exact"
  system = make_system(name)
  tr = make_transcription(system, optz, int(intervals), int(cpi), meth, quad)
  t = time.time()
  r = solve_transcription(tr, IpmOptions(verbose=True, hessian=hessian))
  print("status", r["status"], "iters", r["iters"], "f", r["f"], "time", time.time() - t)
