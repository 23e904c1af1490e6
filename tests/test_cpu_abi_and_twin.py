"""CPU-side tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/myriad_b200.h declares,
descriptor/size logic, error behaviour, and the HOST TWIN of the CUDA templates (same source compiled for the host,
csrc/common.cuh) against the reference fixtures.  No CUDA compute happens here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.cases import CASES, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLLOC = sorted(c for c, v in CASES.items() if v[1] == "COLLOCATION")


@pytest.fixture(scope="session")
def ML():
  from myriad_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    from myriad_b200 import build
    build.build(verbose=False)
  _lib.lib()
  return _lib


def _desc(ML, case):
  # the descriptor the product builds for this case (system parameters, terminal-cost flag; for a NodeSystem the MLP
  # weights are passed as a HOST pointer to the host twin)
  from tests.cases import product_transcription
  return product_transcription(case).desc(device="host")


def p(a):
  return a.ctypes.data_as(C.c_void_p)


def test_library_exports_every_declared_symbol(ML):
  hdr = open(os.path.join(ROOT, "include", "myriad_b200.h")).read()
  declared = set(re.findall(r"\b(myr_[a-z_]+)\s*\(", hdr))
  assert declared == set(ML.EXPORTS), declared ^ set(ML.EXPORTS)
  lib = ML.lib()
  for name in declared:
    assert getattr(lib, name) is not None
  assert lib.myr_abi_version() == ML.ABI_VERSION


@pytest.mark.parametrize("case", sorted(CASES))
def test_problem_sizes_match_reference(ML, case):
  fx = load(case)
  s = ML.problem_sizes(_desc(ML, case))
  assert s.nvars == fx["guess"].shape[0] and s.ncon == fx["con_z"].shape[0]


def test_error_behaviour(ML):
  d = ML.make_desc("CARTPOLE", 7, "HEUN", 10)  # unknown optimizer enum: the reference raises KeyError
  with pytest.raises(KeyError):
    ML.problem_sizes(d)
  d = ML.make_desc("CARTPOLE", ML.OPT_TRAPEZOIDAL, "HEUN", 10)
  d.system_id = 99
  with pytest.raises(KeyError):
    ML.problem_sizes(d)
  d = ML.make_desc("CARTPOLE", ML.OPT_TRAPEZOIDAL, "HEUN", 0)
  with pytest.raises(KeyError):
    ML.problem_sizes(d)
  # workspace too small is reported, not a crash
  d = ML.make_desc("CARTPOLE", ML.OPT_TRAPEZOIDAL, "HEUN", 10)
  s = ML.problem_sizes(d)
  z = np.zeros((1, s.nvars)); o = ML.MyrIpmOpts()
  a = [np.zeros((1, s.nvars)) for _ in range(5)] + [np.zeros((1, s.ncon))]
  sc = [np.zeros(1) for _ in range(3)]; st = np.zeros(1, np.int32); it = np.zeros(1, np.int32); ws = np.zeros(8)
  rc = ML.lib().myr_host_ipm_solve(C.byref(d), C.byref(o), 1, p(z), p(a[0]), p(a[1]), p(a[2]), p(a[5]), p(a[3]), p(a[4]), p(sc[0]), p(sc[1]),
                                   p(sc[2]), p(st), p(it), p(ws), ws.size)
  assert rc == -4 and b"workspace" in ML.lib().myr_last_error()


def _host_eval(ML, d, s, z, lam=None):
  B = z.shape[0]
  f = np.zeros(B); grad = np.zeros((B, s.nvars)); c = np.zeros((B, s.ncon)); J = np.zeros((B, s.jac_block_doubles))
  H = np.zeros((B, s.hess_block_doubles)) if lam is not None else None
  ML.check(ML.lib().myr_host_eval(C.byref(d), B, p(z), p(lam) if lam is not None else None, p(f), p(grad), p(c), p(J),
                                  p(H) if H is not None else None))
  return f, grad, c, J, H


def _dense_J_shooting(s, J, K, cpi, meth):
  n, m = s.n, s.m
  M = (2 if meth == "RK4" else 1) * cpi
  ncol = n + (M + 1) * m
  Jb = J.reshape(K, n + 1, ncol)
  Jd = np.zeros((s.ncon, s.nvars))
  ubase = (K + 1) * n
  for k in range(K):
    Jd[k * n:(k + 1) * n, k * n:(k + 1) * n] += Jb[k, :n, :n]
    Jd[k * n:(k + 1) * n, ubase + k * M * m: ubase + (k * M + M + 1) * m] += Jb[k, :n, n:]
    Jd[k * n:(k + 1) * n, (k + 1) * n:(k + 2) * n] -= np.eye(n)
  return Jd


def _dense_J(s, J, quad):
  n, m, Q = s.n, s.m, s.nodes
  zidx = lambda q, i: q * n + i if i < n else Q * n + q * m + (i - n)
  Jd = np.zeros((s.ncon, s.nvars))
  Jb = J.reshape(s.stages, s.stage_nodes, s.nc, s.nw)
  for j in range(s.stages):
    for k in range(s.stage_nodes):
      q = (j + k) if quad == "TRAPEZOIDAL" else 2 * j + k
      for r in range(s.nc):
        ci = j * n + r if (quad == "TRAPEZOIDAL" or r < n) else s.stages * n + j * n + (r - n)
        for i in range(s.nw):
          Jd[ci, zidx(q, i)] += Jb[j, k, r, i]
  return Jd


@pytest.mark.parametrize("case", sorted(CASES))
def test_host_twin_k1_matches_reference_fixture(ML, case):
  """f, c, grad f and the (block) Jacobian of all three transcriptions x four integrators vs the reference's closures."""
  fx = load(case)
  d = _desc(ML, case)
  s = ML.problem_sizes(d)
  sysname, opt, quad, meth, K, cpi = CASES[case]
  f, grad, c, J, _ = _host_eval(ML, d, s, np.ascontiguousarray(np.stack([fx["z"], fx["guess"]])))
  np.testing.assert_allclose(f, [fx["obj_z"], fx["obj_guess"]], rtol=1e-12, atol=1e-14)
  np.testing.assert_allclose(c, np.stack([fx["con_z"], fx["con_guess"]]), rtol=5e-12, atol=1e-13)  # pow() in TUMOUR: a few ulps
  np.testing.assert_allclose(grad[0], fx["grad_z"], rtol=1e-11, atol=1e-13)
  Jd = _dense_J_shooting(s, J[0], K, cpi, meth) if opt == "SHOOTING" else _dense_J(s, J[0], quad)
  np.testing.assert_allclose(Jd, fx["jac_z"], rtol=1e-11, atol=1e-13)


def _host_ipm(ML, d, s, z0, lb, ub, max_iter=1000):
  B = z0.shape[0]
  out = dict(z=np.zeros((B, s.nvars)), lam=np.zeros((B, s.ncon)), zL=np.zeros((B, s.nvars)), zU=np.zeros((B, s.nvars)),
             obj=np.zeros(B), kkt=np.zeros(B), cinf=np.zeros(B), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32))
  ws = np.zeros(ML.workspace_doubles(s, B))
  o = ML.MyrIpmOpts(); o.max_iter = max_iter
  ML.check(ML.lib().myr_host_ipm_solve(C.byref(d), C.byref(o), B, p(z0), p(lb), p(ub), p(out["z"]), p(out["lam"]), p(out["zL"]), p(out["zU"]),
                                       p(out["obj"]), p(out["kkt"]), p(out["cinf"]), p(out["status"]), p(out["iters"]), p(ws), ws.size))
  return out


@pytest.mark.parametrize("case", [c for c in sorted(CASES) if "sol_cost" in load(c)])
def test_host_twin_ipm_matches_reference_solve(ML, case):
  fx = load(case)
  d = _desc(ML, case)
  s = ML.problem_sizes(d)
  out = _host_ipm(ML, d, s, np.ascontiguousarray(fx["guess"][None]), np.ascontiguousarray(fx["bounds"][None, :, 0]),
                  np.ascontiguousarray(fx["bounds"][None, :, 1]))
  assert out["status"][0] == 0 and out["cinf"][0] <= 1e-8
  ref = float(fx["sol_cost"])
  # SciPy's default ftol=1e-6 leaves SLSQP up to ~3e-5 short of the optimum on the flat SIMPLECASE objectives
  assert abs(out["obj"][0] - ref) <= 5e-5 * max(1.0, abs(ref))
  # the IPM is never worse than SLSQP beyond what SLSQP's own infeasibility buys it (first-order: |lam|_1 * |c_ref|_inf)
  slack = float(np.abs(out["lam"][0]).sum() * fx["sol_con_inf"])
  assert out["obj"][0] <= ref + 1e-7 * max(1.0, abs(ref)) + slack
  # KKT conditions re-checked independently with the oracle's derivatives
  from oracle import nlp
  from oracle.systems import make_system
  from oracle.transcription import make_transcription
  sysname, opt, quad, meth, intervals, cpi = CASES[case]
  tr = make_transcription(make_system(sysname), opt, intervals, cpi, meth, quad)
  z, lam = out["z"][0], out["lam"][0]
  rd = nlp.objective_grad(tr, z) + nlp.constraints_jac(tr, z).T @ lam - out["zL"][0] + out["zU"][0]
  free = fx["bounds"][:, 0] != fx["bounds"][:, 1]
  assert np.abs(rd[free]).max() <= 1e-6 * max(1.0, np.abs(lam).max())
  assert np.abs(tr.constraints(z)).max() <= 1e-8


def test_host_twin_kkt_matches_dense_numpy(ML):
  case = "s_cartpole_trap_10"
  fx = load(case)
  d = _desc(ML, case)
  s = ML.problem_sizes(d)
  rng = np.random.default_rng(3)
  lam = rng.standard_normal((1, s.ncon))
  z = np.ascontiguousarray(fx["z"][None])
  f, grad, c, J, H = _host_eval(ML, d, s, z, lam)
  fixed = fx["bounds"][:, 0] == fx["bounds"][:, 1]
  sigma = rng.uniform(0.1, 2.0, s.nvars)
  sig_in = np.where(fixed, np.inf, sigma)[None].copy()
  rz = rng.standard_normal((1, s.nvars)); rc = rng.standard_normal((1, s.ncon))
  dz = np.zeros((1, s.nvars)); dl = np.zeros((1, s.ncon)); ok = np.zeros(1, np.int32); ws = np.zeros(ML.workspace_doubles(s, 1))
  for delta_w in (0.0, 5.0):
    ML.check(ML.lib().myr_host_kkt_solve(C.byref(d), 1, p(H), p(J), p(sig_in), p(rz), p(rc), delta_w, 0.0, p(dz), p(dl), p(ok), p(ws), ws.size))
    n, m, Q, nw = s.n, s.m, s.nodes, s.nw
    zidx = np.array([[q * n + i if i < n else Q * n + q * m + (i - n) for i in range(nw)] for q in range(Q)])
    Hd = np.zeros((s.nvars, s.nvars)); iu = np.triu_indices(nw)
    for q in range(Q):
      blk = np.zeros((nw, nw)); blk[iu] = H[0].reshape(Q, -1)[q]; blk = blk + blk.T - np.diag(np.diag(blk))
      Hd[np.ix_(zidx[q], zidx[q])] = blk
    Jd = _dense_J(s, J[0], "TRAPEZOIDAL")
    fr = np.where(~fixed)[0]
    K = np.block([[Hd[np.ix_(fr, fr)] + np.diag(sigma[fr] + delta_w), Jd[:, fr].T], [Jd[:, fr], np.zeros((s.ncon, s.ncon))]])
    ev = np.linalg.eigvalsh(K)
    assert int(ok[0]) == int((ev < 0).sum() == s.ncon and np.abs(ev).min() > 1e-10)
    sol = np.linalg.solve(K, -np.concatenate([rz[0][fr], rc[0]]))
    ref = np.zeros(s.nvars); ref[fr] = sol[:len(fr)]
    scale = max(1.0, np.abs(sol).max())
    np.testing.assert_allclose(dz[0], ref, atol=1e-9 * scale)
    np.testing.assert_allclose(dl[0], sol[len(fr):], atol=1e-9 * scale)


@pytest.mark.parametrize("case", sorted(CASES))
def test_host_twin_rollout_matches_reference_fixture(ML, case):
  fx = load(case)
  if "rollout_states" not in fx or not np.isfinite(fx["rollout_states"]).all():
    pytest.skip("no finite reference rollout for this combination")
  d = _desc(ML, case)
  s = ML.problem_sizes(d)
  sysname, opt, quad, meth, intervals, cpi = CASES[case]
  from oracle.systems import make_system
  x0 = np.ascontiguousarray(np.asarray(make_system(sysname).x_0, dtype=np.float64)[None])
  u = np.ascontiguousarray(fx["z"][s.nx_nodes * s.n:].reshape(1, s.nu_nodes, s.m))
  steps = intervals * (cpi if opt == "SHOOTING" else 1)
  xs = np.zeros((1, steps + 1, s.n)); cost = np.zeros(1)
  ML.check(ML.lib().myr_host_rollout_cost(C.byref(d), 1, s.nu_nodes, p(u), p(x0), p(xs), p(cost)))
  np.testing.assert_allclose(xs[0], fx["rollout_states"], rtol=1e-12, atol=1e-13)
  np.testing.assert_allclose(cost[0], float(fx["rollout_cost"]), rtol=1e-12)
