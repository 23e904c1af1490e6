"""NumPy model of the block-structured KKT solve used by the CUDA K2 kernel (development tool).

Collocation KKT (trapezoid):   [ H  J^T ] [dz ]   [ -rb ]
                               [ J  -dc ] [dlam] = [ -c  ]
H block-diagonal per node (possibly INDEFINITE), J block-bidiagonal (row k: G_k on node k, F_k on
node k+1).  Schur complement S = J H^-1 J^T (+dc) is block tridiagonal; block cyclic reduction
factorises it.  Inertia(K) = Inertia(H) + Inertia(-S)  (Haynsworth), and Inertia(S) is the sum of the
inertias of the CR pivot blocks, so the IPOPT-style inertia test needs no extra factorisation.
"""
import numpy as np


def sym_inertia(A, tol=1e-14):
  ev = np.linalg.eigvalsh(A)
  s = np.abs(ev).max() if ev.size else 0.0
  return int((ev > tol * s).sum()), int((ev < -tol * s).sum()), int((np.abs(ev) <= tol * s).sum())


def block_cr_solve(D, U, b):
  """Solve block-tridiagonal symmetric system: D[i] x_i + U[i] x_{i+1} + U[i-1]^T x_{i-1} = b[i].
  D: (N,bs,bs), U: (N-1,bs,bs), b: (N,bs).  Returns x and (npos, nneg, nzero) summed over pivots."""
  N = D.shape[0]
  D = D.copy(); b = b.copy()
  idx = list(range(N))
  # generic recursive odd-even elimination on an index list with coupling dict
  Ud = {(i, i + 1): U[i].copy() for i in range(N - 1)}
  levels = []
  inert = np.zeros(3, dtype=int)
  active = idx
  while len(active) > 1:
    elim = active[1::2]
    keep = active[0::2]
    pos = {a: p for p, a in enumerate(active)}
    rec = []
    newU = {}
    for e in elim:
      p = pos[e]
      l = active[p - 1]
      r = active[p + 1] if p + 1 < len(active) else None
      inert += sym_inertia(D[e])
      Dinv = np.linalg.inv(D[e])
      Ul = Ud[(l, e)]  # coupling l -> e  (row l, col e)
      Ur = Ud[(e, r)] if r is not None else None
      rec.append((e, l, r, Dinv, Ul, Ur))
    # gather updates
    Dn = {k: D[k].copy() for k in keep}
    bn = {k: b[k].copy() for k in keep}
    for (e, l, r, Dinv, Ul, Ur) in rec:
      Dn[l] -= Ul @ Dinv @ Ul.T
      bn[l] -= Ul @ Dinv @ b[e]
      if r is not None:
        Dn[r] -= Ur.T @ Dinv @ Ur
        bn[r] -= Ur.T @ Dinv @ b[e]
        newU[(l, r)] = -Ul @ Dinv @ Ur
    for k in keep:
      D[k] = Dn[k]; b[k] = bn[k]
    Ud.update(newU)
    levels.append(rec)
    active = keep
  root = active[0]
  inert += sym_inertia(D[root])
  x = np.zeros_like(b)
  x[root] = np.linalg.solve(D[root], b[root])
  for rec in reversed(levels):
    for (e, l, r, Dinv, Ul, Ur) in rec:
      rhs = b[e] - Ul.T @ x[l]
      if r is not None:
        rhs = rhs - Ur @ x[r]
      x[e] = Dinv @ rhs
  return x, tuple(inert)


def schur_cr_kkt_solve(Hb, G, F, rb, c, free_b, delta_c=0.0):
  """Hb: (L,nw,nw) node Hessian blocks (already + Sigma + delta); G,F: (N,n,nw); rb: (L,nw); c: (N,n);
  free_b: (L,nw) bool.  Returns dz (L,nw), dlam (N,n), inertia of K as (npos,nneg,nzero)."""
  L, nw, _ = Hb.shape
  N, n, _ = G.shape
  Hm = Hb.copy(); Gm = G.copy(); Fm = F.copy(); rm = rb.copy()
  inertH = np.zeros(3, dtype=int)
  Hinv = np.zeros_like(Hm)
  for k in range(L):
    fr = free_b[k]
    sub = Hm[k][np.ix_(fr, fr)]
    inertH += sym_inertia(sub)
    inv = np.zeros((nw, nw))
    if fr.any():
      inv[np.ix_(fr, fr)] = np.linalg.inv(sub)
    Hinv[k] = inv  # zero rows/cols for fixed variables == masking J columns, dz_fixed = 0
  D = np.zeros((N, n, n)); U = np.zeros((N - 1, n, n)); b = np.zeros((N, n))
  for k in range(N):
    D[k] = G[k] @ Hinv[k] @ G[k].T + F[k] @ Hinv[k + 1] @ F[k].T + delta_c * np.eye(n)
    b[k] = c[k] - G[k] @ Hinv[k] @ rb[k] - F[k] @ Hinv[k + 1] @ rb[k + 1]
    if k + 1 < N:
      U[k] = F[k] @ Hinv[k + 1] @ G[k + 1].T
  dlam, inertS = block_cr_solve(D, U, b)
  dz = np.zeros((L, nw))
  for k in range(L):
    v = rb[k].copy()
    if k < N:
      v += G[k].T @ dlam[k]
    if k > 0:
      v += F[k - 1].T @ dlam[k - 1]
    dz[k] = -Hinv[k] @ v
  # inertia of K: In(H) + In(-S)
  npos = inertH[0] + inertS[1]; nneg = inertH[1] + inertS[0]; nzero = inertH[2] + inertS[2]
  return dz, dlam, (npos, nneg, nzero)


def blocks_from_dense(tr, W, J, vecs):
  """Split dense W (nv,nv), J (ncon,nv) and per-variable vectors into node blocks for a trapezoid transcription."""
  n, m, L = tr.n, tr.m, tr.nx_nodes
  N = L - 1
  nw = n + m
  ix = [np.concatenate([np.arange(k * n, (k + 1) * n), L * n + np.arange(k * m, (k + 1) * m)]) for k in range(L)]
  Hb = np.stack([W[np.ix_(ix[k], ix[k])] for k in range(L)])
  G = np.stack([J[k * n:(k + 1) * n][:, ix[k]] for k in range(N)])
  F = np.stack([J[k * n:(k + 1) * n][:, ix[k + 1]] for k in range(N)])
  out = [np.stack([v[ix[k]] for k in range(L)]) for v in vecs]
  return ix, Hb, G, F, out


if __name__ == "__main__":
  import sys
  sys.path.insert(0, ".")
  from oracle.systems import make_system
  from oracle.transcription import make_transcription
  from tools.ipm_proto import IpmOptions, solve_transcription
  N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
  tr = make_transcription(make_system("CARTPOLE"), "COLLOCATION", N, 1, "HEUN", "TRAPEZOIDAL")
  cap = []
  r = solve_transcription(tr, IpmOptions(capture=cap))
  print("ipm", r["status"], r["iters"], r["f"])
  for it, d in enumerate(cap):
    W, J, Sig, rb, c, free, delta = d["W"], d["J"], d["Sig"], d["rb"], d["c"], d["free"], d["delta"]
    nv = W.shape[0]; nc = J.shape[0]
    Hfull = W + np.diag(Sig + delta)
    idx = np.where(free)[0]
    K = np.block([[Hfull[np.ix_(idx, idx)], J[:, idx].T], [J[:, idx], np.zeros((nc, nc))]])
    sol = np.linalg.solve(K, -np.concatenate([rb[idx], c]))
    dz_ref = np.zeros(nv); dz_ref[idx] = sol[:len(idx)]; dl_ref = sol[len(idx):]
    ix, Hb, G, F, (rbb, freeb) = blocks_from_dense(tr, Hfull, J, [rb, free])
    dz_b, dl_b, inert = schur_cr_kkt_solve(Hb, G, F, rbb, c.reshape(-1, tr.n), freeb.astype(bool))
    dz = np.zeros(nv)
    for k in range(len(ix)):
      dz[ix[k]] = dz_b[k]
    e1 = np.abs(dz - dz_ref).max() / max(1e-300, np.abs(dz_ref).max())
    e2 = np.abs(dl_b.ravel() - dl_ref).max() / max(1e-300, np.abs(dl_ref).max())
    evK = np.linalg.eigvalsh(K)
    print(f"it {it:2d} delta={delta:.1e} err dz={e1:.2e} dlam={e2:.2e} inertia CR={inert} true=({(evK>0).sum()},{(evK<0).sum()}) "
          f"cond(K)={np.abs(evK).max()/np.abs(evK).min():.1e}")
    # inertia detection at other deltas (incl. the rejected delta=0)
    for dtest in (0.0, 1e-4, 1e-2):
      Hf = W + np.diag(Sig + dtest)
      Kt = np.block([[Hf[np.ix_(idx, idx)], J[:, idx].T], [J[:, idx], np.zeros((nc, nc))]])
      ev = np.linalg.eigvalsh(Kt)
      _, Hb2, G2, F2, (rbb2, freeb2) = blocks_from_dense(tr, Hf, J, [rb, free])
      _, _, inert2 = schur_cr_kkt_solve(Hb2, G2, F2, rbb2, c.reshape(-1, tr.n), freeb2.astype(bool))
      nnegH = sum(sym_inertia(Hb2[k][np.ix_(freeb2[k].astype(bool), freeb2[k].astype(bool))])[1] for k in range(Hb2.shape[0]))
      flag = "OK " if (int(inert2[0]), int(inert2[1])) == (int((ev > 0).sum()), int((ev < 0).sum())) else "MISMATCH"
      print(f"      delta={dtest:.0e}: CR inertia=({inert2[0]},{inert2[1]},{inert2[2]}) true=({(ev>0).sum()},{(ev<0).sum()}) n-(H)={nnegH} {flag}")
