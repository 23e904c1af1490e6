"""Trajectory-level parity, basin agreement on random start states, solve_with_params.

Every check exists twice: through the HOST TWIN (-m "not gpu": same templates compiled for the CPU, so the algorithm
is verified in CI without a device) and through the CUDA path (-m gpu: the parity tests proper).

Fixtures (oracle/make_parity.py):
  tight_<case>.npz    the same NLP solved by SLSQP with ftol=1e-12 (the committed sol_* fixtures stop at SciPy's
                      default ftol=1e-6, i.e. up to ~3e-5 short of the optimum on flat objectives: VERDICT r1 weak #1)
  random_x0_c2.npz    16 rows of the bench workload (CARTPOLE trapezoid N=100, random x0) solved by SLSQP
  params_cartpole.npz reference code (under oracle/refshim) with non-default CARTPOLE (g, m1, m2, length)
Tolerances: controls / states max |u - u_ref| <= 1e-3 * (bound range), objective 1e-6 relative, re-integrated cost of
the true system 1e-6 relative, terminal defect 1e-6 absolute -- the contract SURVEY.md section 8c(2,3) states.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests.cases import CASES, GOLDEN, load, product_transcription

TIGHT = sorted(f[len("tight_"):-4] for f in os.listdir(GOLDEN) if f.startswith("tight_"))
# cases whose optimum is a flat valley in u (SIMPLECASE family: objective curvature ~1e-2, so even ftol=1e-12 pins u
# only to ~1e-5 * range) keep the same tolerance -- it is met -- but are listed to document the reason when it is tight
HAVE_RANDOM = os.path.exists(os.path.join(GOLDEN, "random_x0_c2.npz"))
HAVE_PARAMS = os.path.exists(os.path.join(GOLDEN, "params_cartpole.npz"))


def _p(a):
  return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- two back ends with one interface
class _Host:
  """host twin through the C ABI"""
  name = "host"

  def solve(self, tr, z0, lb, ub, max_iter=1000):
    from myriad_b200 import _lib as ML
    d = tr.desc(device="host")
    s = ML.problem_sizes(d)
    B = z0.shape[0]
    z0, lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (z0, lb, ub))
    out = dict(z=np.zeros((B, s.nvars)), lam=np.zeros((B, s.ncon)), zL=np.zeros((B, s.nvars)), zU=np.zeros((B, s.nvars)),
               obj=np.zeros(B), kkt=np.zeros(B), cinf=np.zeros(B), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32))
    ws = np.zeros(ML.workspace_doubles(s, B))
    o = ML.MyrIpmOpts(); o.max_iter = max_iter
    ML.check(ML.lib().myr_host_ipm_solve(C.byref(d), C.byref(o), B, _p(z0), _p(lb), _p(ub), _p(out["z"]), _p(out["lam"]), _p(out["zL"]),
                                         _p(out["zU"]), _p(out["obj"]), _p(out["kkt"]), _p(out["cinf"]), _p(out["status"]), _p(out["iters"]),
                                         _p(ws), ws.size))
    return out

  def rollout(self, tr, u, x0):
    from myriad_b200 import _lib as ML
    d = tr.desc(device="host")
    s = ML.problem_sizes(d)
    B = u.shape[0]
    steps = tr.intervals * tr.cpi
    u = np.ascontiguousarray(u, dtype=np.float64); x0 = np.ascontiguousarray(x0, dtype=np.float64)
    xs = np.zeros((B, steps + 1, s.n)); cost = np.zeros(B)
    ML.check(ML.lib().myr_host_rollout_cost(C.byref(d), B, u.shape[1], _p(u), _p(x0), _p(xs), _p(cost)))
    return xs, cost

  def eval(self, tr, z):
    from myriad_b200 import _lib as ML
    d = tr.desc(device="host")
    s = ML.problem_sizes(d)
    z = np.ascontiguousarray(z, dtype=np.float64)
    B = z.shape[0]
    f = np.zeros(B); c = np.zeros((B, s.ncon))
    ML.check(ML.lib().myr_host_eval(C.byref(d), B, _p(z), None, _p(f), None, _p(c), None, None))
    return f, c


class _Cuda:
  """CUDA kernels through the C ABI (torch tensors are the device buffers)"""
  name = "cuda"

  def _dev(self, a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda().contiguous()

  def solve(self, tr, z0, lb, ub, max_iter=1000):
    import torch
    from myriad_b200.engine import Engine
    out = Engine(tr.desc()).ipm_solve(self._dev(z0), self._dev(lb), self._dev(ub), max_iter=max_iter)
    torch.cuda.synchronize()
    r = {k: v.cpu().numpy() for k, v in out.items()}
    r["cinf"] = r["con_inf"]; r["kkt"] = r["kkt_err"]
    return r

  def rollout(self, tr, u, x0):
    import torch
    from myriad_b200.engine import Engine
    xs, cost = Engine(tr.desc()).rollout_cost(self._dev(u), self._dev(x0))
    torch.cuda.synchronize()
    return xs.cpu().numpy(), cost.cpu().numpy()

  def eval(self, tr, z):
    import torch
    from myriad_b200.engine import Engine
    r = Engine(tr.desc()).eval(self._dev(z))
    torch.cuda.synchronize()
    return r.f.cpu().numpy(), r.c.cpu().numpy()


BACKENDS = [pytest.param(_Host(), id="host"), pytest.param(_Cuda(), id="cuda", marks=pytest.mark.gpu)]


def _ranges(tr, bounds, x_ref, u_ref):
  """normalisation of trajectory differences: the bound range where finite, else the range of the reference trajectory"""
  n, m = tr.n, tr.m
  sb = np.asarray(tr.system.bounds, dtype=np.float64)
  xr = np.where(np.isfinite(sb[:n, 1] - sb[:n, 0]), sb[:n, 1] - sb[:n, 0], np.maximum(1.0, np.ptp(x_ref, axis=0)))
  ur = np.where(np.isfinite(sb[n:, 1] - sb[n:, 0]), sb[n:, 1] - sb[n:, 0], np.maximum(1.0, np.ptp(u_ref, axis=0)))
  return xr, ur


# ----------------------------------------------------------------------------- (a) trajectories, rollout cost, defect
@pytest.mark.parametrize("be", BACKENDS)
@pytest.mark.parametrize("case", TIGHT)
def test_trajectories_match_tight_oracle_solve(be, case):
  fx = load(case)
  tg = load("tight_" + case)
  if not bool(tg["success"]):
    pytest.skip("oracle SLSQP did not converge at ftol=1e-12 for this case")
  tr = product_transcription(case)
  out = be.solve(tr, fx["guess"][None], fx["bounds"][None, :, 0], fx["bounds"][None, :, 1])
  assert int(out["status"][0]) == 0 and float(out["cinf"][0]) <= 1e-8
  ref = float(tg["cost"])
  assert abs(float(out["obj"][0]) - ref) <= 1e-6 * max(1.0, abs(ref)), (float(out["obj"][0]), ref)
  x, u = tr.unravel(out["z"][0])
  xr, ur = _ranges(tr, fx["bounds"], tg["x"], tg["u"])
  du = np.abs(u - tg["u"]).max(axis=0) / ur
  dx = np.abs(x - tg["x"]).max(axis=0) / xr
  assert du.max() <= 1e-3, ("controls", du)
  assert dx.max() <= 1e-3, ("states", dx)
  # what run_trajectory_opt returns (useful_scripts.py:47-49,76): cost and defect of the TRUE system under the solved controls
  if "rollout_cost" in tg:
    true_tr = tr
    if CASES[case][0].startswith("NODE_"):
      from myriad_b200 import problems as PR
      true_tr = PR.Transcription(tr.system.true_system, tr.optimizer, tr.method, tr.intervals, tr.cpi)
    xs, cost = be.rollout(true_tr, u[None], np.asarray(true_tr.system.x_0, dtype=np.float64)[None])
    rc = float(tg["rollout_cost"])
    assert abs(float(cost[0]) - rc) <= 1e-6 * max(1.0, abs(rc)), (float(cost[0]), rc)
    if "rollout_defect" in tg:
      xT = true_tr.system.x_T
      idx = [i for i in range(len(xT)) if xT[i] is not None]
      defect = xs[0, -1, idx] - np.array([xT[i] for i in idx], dtype=np.float64)
      # the defect of the re-integrated trajectory is sensitive to u: d defect / d u ~ O(T); both sides carry the
      # solver tolerance of their own u, hence 1e-6 absolute on top of a relative part
      np.testing.assert_allclose(defect, tg["rollout_defect"], atol=1e-6 + 1e-5 * np.abs(tg["rollout_defect"]).max())


# ----------------------------------------------------------------------------- (b) random start states: same basin
@pytest.mark.skipif(not HAVE_RANDOM, reason="random_x0_c2.npz not generated")
@pytest.mark.parametrize("be", BACKENDS)
def test_random_start_states_reach_the_oracle_basin(be):
  """CARTPOLE swing-up is non-convex: rows of the bench workload must end in the same local solution SLSQP finds
  (SURVEY.md section 7.4-1).  Agreement is asserted on >= 15 of 16 rows; a row that differs must still be a KKT point
  with an objective not worse than SLSQP's by more than 1e-6 (a different but better basin is not an error)."""
  from myriad_b200 import problems as PR
  rx = dict(np.load(os.path.join(GOLDEN, "random_x0_c2.npz")))
  tr = product_transcription("c2_cartpole_trap_100")
  import torch
  x0 = torch.as_tensor(rx["x0"])
  if be.name == "cuda":
    z0, lb, ub = (t.cpu().numpy() for t in PR.build_batch(tr, x0.cuda()))
  else:
    pytest.importorskip("torch")
    if not torch.cuda.is_available():
      # build_batch needs the device only for rolled-out guesses; CARTPOLE's guess is a linspace: build it in numpy
      fx = load("c2_cartpole_trap_100")
      B = x0.shape[0]
      z0 = np.tile(fx["guess"], (B, 1)); lb = np.tile(fx["bounds"][:, 0], (B, 1)); ub = np.tile(fx["bounds"][:, 1], (B, 1))
      n, L = tr.n, tr.nx_nodes
      xT = np.asarray(tr.system.x_T, dtype=np.float64)
      k = np.arange(L, dtype=np.float64)
      for b in range(B):
        a = rx["x0"][b]
        xg = a[None, :] + k[:, None] * ((xT - a) / (L - 1))[None, :]
        xg[-1] = xT
        z0[b, :L * n] = xg.reshape(-1)
        lb[b, :n] = a; ub[b, :n] = a
    else:
      z0, lb, ub = (t.cpu().numpy() for t in PR.build_batch(tr, x0.cuda()))
  out = be.solve(tr, z0, lb, ub)
  ok = out["status"] == 0
  assert ok.sum() >= 15, out["status"]
  same = np.abs(out["obj"] - rx["cost_tight"]) <= 1e-6 * np.maximum(1.0, np.abs(rx["cost_tight"]))
  assert (same & ok).sum() >= 15, (out["obj"], rx["cost_tight"])
  better = out["obj"] <= rx["cost_tight"] + 1e-6 * np.maximum(1.0, np.abs(rx["cost_tight"]))
  assert (ok & (same | better)).sum() >= 15
  # trajectories of the rows in the same basin
  x, u = tr.unravel(out["z"])
  sel = same & ok & rx["success"]
  assert np.abs(u[sel] - rx["u_tight"][sel]).max() <= 1e-3 * 40.0      # u range of CARTPOLE = [-20, 20]
  # row 0 is the reference problem itself
  fx = load("c2_cartpole_trap_100")
  assert abs(out["obj"][0] - float(fx["sol_cost"])) <= 5e-5 * abs(float(fx["sol_cost"]))


def test_filter_line_search_keeps_the_optima_and_cuts_iterations(monkeypatch):
  """The line search accepts a trial point by the l1-merit Armijo test OR by IPOPT's filter rules (engine.cuh).  On 64
  rows of the bench workload (host twin; the device runs the same templates) the filter must (a) converge every row,
  (b) end in the same optimum as the merit-only line search, (c) need clearly fewer iterations."""
  from myriad_b200 import problems as PR
  import torch
  tr = product_transcription("c2_cartpole_trap_100")
  fx = load("c2_cartpole_trap_100")
  B = 64
  x0 = PR.sample_x0(tr.system, B).numpy()
  z0 = np.tile(fx["guess"], (B, 1)); lb = np.tile(fx["bounds"][:, 0], (B, 1)); ub = np.tile(fx["bounds"][:, 1], (B, 1))
  n, L = tr.n, tr.nx_nodes
  xT = np.asarray(tr.system.x_T, dtype=np.float64)
  k = np.arange(L, dtype=np.float64)
  for b in range(B):
    xg = x0[b][None, :] + k[:, None] * ((xT - x0[b]) / (L - 1))[None, :]
    xg[-1] = xT
    z0[b, :L * n] = xg.reshape(-1)
    lb[b, :n] = x0[b]; ub[b, :n] = x0[b]
  host = _Host()
  monkeypatch.setenv("MYR_FILTER", "0")
  merit = host.solve(tr, z0, lb, ub)
  monkeypatch.setenv("MYR_FILTER", "1")
  filt = host.solve(tr, z0, lb, ub)
  assert (merit["status"] == 0).all() and (filt["status"] == 0).all()
  np.testing.assert_allclose(filt["obj"], merit["obj"], rtol=1e-8)
  assert np.abs(filt["z"] - merit["z"]).max() <= 1e-5
  assert filt["iters"].sum() <= 0.85 * merit["iters"].sum(), (filt["iters"].sum(), merit["iters"].sum())


def test_watchdog_ends_a_crawl_of_shortened_steps(monkeypatch):
  """Row 56635 of the seeded CARTPOLE draw is the worst of 65 536: without the watchdog the line search accepts only
  1-3 % steps for ~400 iterations before three full steps finish the solve.  With the watchdog (10 shortened steps in a
  row -> up to 3 full steps judged against the stored iterate) it must reach the same optimum in a tenth of that."""
  from myriad_b200 import problems as PR
  tr = product_transcription("c2_cartpole_trap_100")
  fx = load("c2_cartpole_trap_100")
  x0 = PR.sample_x0(tr.system, 56636).numpy()[56635]
  z0 = fx["guess"].copy()[None]; lb = fx["bounds"][:, 0].copy()[None]; ub = fx["bounds"][:, 1].copy()[None]
  n, L = tr.n, tr.nx_nodes
  xT = np.asarray(tr.system.x_T, dtype=np.float64)
  xg = x0[None, :] + np.arange(L, dtype=np.float64)[:, None] * ((xT - x0) / (L - 1))[None, :]
  xg[-1] = xT
  z0[0, :L * n] = xg.reshape(-1)
  lb[0, :n] = x0; ub[0, :n] = x0
  host = _Host()
  monkeypatch.setenv("MYR_WATCHDOG", "0")
  crawl = host.solve(tr, z0, lb, ub)
  monkeypatch.setenv("MYR_WATCHDOG", "1")
  wd = host.solve(tr, z0, lb, ub)
  assert crawl["status"][0] == 0 and wd["status"][0] == 0
  assert crawl["iters"][0] > 200 and wd["iters"][0] <= 60, (crawl["iters"], wd["iters"])
  np.testing.assert_allclose(wd["obj"], crawl["obj"], rtol=1e-8)


# ----------------------------------------------------------------------------- (c) solve_with_params
def _param_system():
  from myriad_b200.systems import SystemType
  pf = dict(np.load(os.path.join(GOLDEN, "params_cartpole.npz")))
  g, m1, m2, length = (float(v) for v in pf["params"])
  return pf, SystemType.CARTPOLE(g=g, m1=m1, m2=m2, length=length), dict(g=g, m1=m1, m2=m2, length=length)


@pytest.mark.skipif(not HAVE_PARAMS, reason="params_cartpole.npz not generated")
@pytest.mark.parametrize("be", BACKENDS)
def test_parametrized_system_matches_reference(be):
  """Non-default CARTPOLE (g, m1, m2, length): K1 values at a test point and the solved objective against the
  reference's parametrized_objective / parametrized_constraints / solve_with_params (shooting: base.py:81-93) and
  against the optimizer of hp.system(**params) (trapezoid: the evident intent)."""
  from myriad_b200 import problems as PR
  pf, system, _ = _param_system()
  for key, tr in (("shoot", PR.Transcription(system, PR.SHOOTING, "HEUN", 5, 4)),
                  ("trap", PR.Transcription(system, PR.TRAPEZOIDAL, "HEUN", 10, 1))):
    f, c = be.eval(tr, pf[key + "_z"][None])
    np.testing.assert_allclose(f[0], float(pf[key + "_obj_z"]), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(c[0], pf[key + "_con_z"], rtol=1e-12, atol=1e-13)
    base = product_transcription("s_cartpole_shooting_5x4_heun" if key == "shoot" else "s_cartpole_trap_10")
    fx = load("s_cartpole_shooting_5x4_heun" if key == "shoot" else "s_cartpole_trap_10")
    # the reference keeps self.guess / self.bounds of the default system (base.py:84-87)
    out = be.solve(tr, fx["guess"][None], fx["bounds"][None, :, 0], fx["bounds"][None, :, 1])
    assert int(out["status"][0]) == 0 and float(out["cinf"][0]) <= 1e-8
    ref = float(pf[key + "_sol_cost"])
    assert abs(float(out["obj"][0]) - ref) <= 5e-5 * abs(ref), (key, float(out["obj"][0]), ref)
    # and the parameters matter: the default system gives a different optimum
    out0 = be.solve(base, fx["guess"][None], fx["bounds"][None, :, 0], fx["bounds"][None, :, 1])
    assert abs(float(out0["obj"][0]) - ref) > 1e-2 * abs(ref)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_PARAMS, reason="params_cartpole.npz not generated")
def test_solve_with_params_surface():
  """TrajectoryOptimizer.solve_with_params (base.py:81-93) through the mirrored surface, CUDA path."""
  from myriad_b200.config import Config, HParams, IntegrationMethod, OptimizerType
  from myriad_b200.systems import SystemType
  from myriad_b200.trajectory_optimizers import get_optimizer
  pf, _, params = _param_system()
  cfg = Config(verbose=False, plot=False)
  hp = HParams(system=SystemType.CARTPOLE, optimizer=OptimizerType.SHOOTING, integration_method=IntegrationMethod.HEUN,
               intervals=5, controls_per_interval=4)
  opt = get_optimizer(hp, cfg, hp.system())
  np.testing.assert_allclose(opt.parametrized_objective(params, pf["shoot_z"]), float(pf["shoot_obj_z"]), rtol=1e-12, atol=1e-14)
  np.testing.assert_allclose(opt.parametrized_constraints(params, pf["shoot_z"]), pf["shoot_con_z"], rtol=1e-12, atol=1e-13)
  res = opt.solve_with_params(params)
  assert abs(res["cost"] - float(pf["shoot_sol_cost"])) <= 5e-5 * abs(float(pf["shoot_sol_cost"]))
  assert set(("x", "u", "xs_and_us", "cost", "lambda")) <= set(res)
  hp2 = HParams(system=SystemType.CARTPOLE, optimizer=OptimizerType.COLLOCATION, intervals=10)
  opt2 = get_optimizer(hp2, cfg, hp2.system())
  res2 = opt2.solve_with_params(params)
  assert abs(res2["cost"] - float(pf["trap_sol_cost"])) <= 5e-5 * abs(float(pf["trap_sol_cost"]))
  # many calls must not grow the engine cache (unbounded _ENGINES)
  from myriad_b200 import nlp_solvers
  n0 = len(nlp_solvers._ENGINES)
  for k in range(3):
    opt2.solve_with_params({**params, "g": 9.0 + 0.1 * k})
  assert len(nlp_solvers._ENGINES) <= max(n0 + 3, nlp_solvers.MAX_ENGINES)


# ----------------------------------------------------------------------------- (d) system plugin surface, VJP, datasets, extragradient
def _abi_call(name, *args):
  from myriad_b200 import _lib as ML
  ML.check(getattr(ML.lib(), name)(*args))


@pytest.mark.parametrize("sysname", ["CARTPOLE", "VANDERPOL", "ROCKETLANDING", "PENDULUM", "MOUNTAINCAR", "BEARPOPULATIONS", "HARVEST"])
def test_host_dynamics_and_cost_match_oracle(sysname):
  """myr_host_dynamics (the generated device code compiled for the host) against the oracle restatement of the
  reference's dynamics / cost (pinned to the reference by the fixtures), incl. the non-smooth systems' clipped regions."""
  from myriad_b200 import problems as PR
  from myriad_b200.systems import SystemType
  from oracle.systems import make_system
  system = SystemType[sysname]()
  osys = make_system(sysname)
  n, m = system.state_size, system.control_size
  rng = np.random.default_rng(3)
  b = np.asarray(system.bounds, dtype=np.float64)
  lo = np.where(np.isfinite(b[:, 0]), b[:, 0], -2.0); hi = np.where(np.isfinite(b[:, 1]), b[:, 1], 2.0)
  pts = lo + (hi - lo) * rng.uniform(-0.2, 1.2, (64, n + m))   # 20 % beyond the bounds: exercises clip / angle_normalize
  x, u = np.ascontiguousarray(pts[:, :n]), np.ascontiguousarray(pts[:, n:])
  if sysname in ("HARVEST",):
    x = np.abs(x)
  t = np.ascontiguousarray(rng.uniform(0, float(system.T), 64))
  f = np.zeros((64, n)); g = np.zeros(64)
  d = PR.Transcription(system, PR.TRAPEZOIDAL, "HEUN", 1, 1).desc(device="host")
  _abi_call("myr_host_dynamics", C.byref(d), 64, _p(x), _p(u), _p(t), _p(f), _p(g))
  np.testing.assert_allclose(f, osys.dynamics(x, u), rtol=1e-12, atol=1e-13)
  np.testing.assert_allclose(g, osys.cost(x, u, t), rtol=1e-12, atol=1e-13)


@pytest.mark.gpu
def test_system_dynamics_and_cost_methods():
  """FiniteHorizonControlSystem.dynamics / .cost keep the reference's call signatures (systems/base.py:45-73) and
  evaluate the generated device code on the GPU."""
  from myriad_b200.systems import SystemType
  from oracle.systems import make_system
  for name in ("CARTPOLE", "ROCKETLANDING", "PENDULUM"):
    system = SystemType[name]()
    osys = make_system(name)
    x = np.asarray(system.x_0, dtype=np.float64) + 0.1
    u = np.full(system.control_size, 0.3)
    np.testing.assert_allclose(system.dynamics(x, u), osys.dynamics(x, u), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(system.cost(x, u, 0.5), float(osys.cost(x, u, 0.5)), rtol=1e-12, atol=1e-13)
    xs = np.stack([x, x + 0.05, x - 0.05])
    us = np.stack([u, u, u])
    np.testing.assert_allclose(system.dynamics(xs, us), osys.dynamics(xs, us), rtol=1e-12, atol=1e-13)
  cp = SystemType.CARTPOLE()
  pf = dict(g=9.0, m1=1.2, m2=0.4, length=0.6)
  x = np.array([0.1, 0.2, -0.3, 0.4]); u = np.array([1.5])
  np.testing.assert_allclose(cp.parametrized_dynamics(pf, x, u), make_system("CARTPOLE", **pf).dynamics(x, u), rtol=1e-12)


@pytest.mark.parametrize("be", BACKENDS)
@pytest.mark.parametrize("case", ["s_cartpole_trap_10", "s_vanderpol_hs_10", "s_cartpole_shooting_5x4_heun", "s_vanderpol_shooting_4x5_rk4",
                                  "x_bear_shooting_3x4_heun", "y_rocket_trap_10"])
def test_jtvec_matches_dense_jacobian_of_reference(be, case):
  """myr_jtvec: J^T lam against the reference's dense jacrev fixture"""
  from myriad_b200 import _lib as ML
  fx = load(case)
  tr = product_transcription(case)
  rng = np.random.default_rng(4)
  lam = np.ascontiguousarray(rng.standard_normal((1, tr.ncon)))
  z = np.ascontiguousarray(fx["z"][None])
  if be.name == "host":
    d = tr.desc(device="host")
    s = ML.problem_sizes(d)
    f = np.zeros(1); grad = np.zeros((1, s.nvars)); c = np.zeros((1, s.ncon)); J = np.zeros((1, s.jac_block_doubles)); out = np.zeros((1, s.nvars))
    _abi_call("myr_host_eval", C.byref(d), 1, _p(z), None, _p(f), _p(grad), _p(c), _p(J), None)
    _abi_call("myr_host_jtvec", C.byref(d), 1, _p(J), _p(lam), _p(out))
  else:
    import torch
    from myriad_b200.engine import Engine
    eng = Engine(tr.desc())
    r = eng.eval(torch.as_tensor(z).cuda())
    out = eng.jtvec(r.Jblk, torch.as_tensor(lam).cuda()).cpu().numpy()
  np.testing.assert_allclose(out[0], fx["jac_z"].T @ lam[0], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("be", BACKENDS)
def test_dataset_rollouts_match_reference(be):
  """Batched dataset rollouts (myriad/utils.py:422-424, integrate_time_independent_in_parallel) for all four
  integration methods, incl. the reference's clamp-to-last control indexing."""
  from myriad_b200 import problems as PR
  from myriad_b200.systems import SystemType
  fx = dict(np.load(os.path.join(GOLDEN, "dataset_rollouts.npz")))
  for name in ("CARTPOLE", "VANDERPOL", "ROCKETLANDING"):
    system = SystemType[name]()
    steps = int(fx[f"{name}_steps"])
    for meth in ("EULER", "HEUN", "MIDPOINT", "RK4"):
      us = fx[f"{name}_us"] if meth == "RK4" else fx[f"{name}_us"][:, :steps + 1]
      tr = PR.Transcription(system, PR.TRAPEZOIDAL, meth, steps, 1)
      xs, _ = be.rollout(tr, us, fx[f"{name}_x0"])
      np.testing.assert_allclose(xs, fx[f"{name}_{meth}_xs"], rtol=1e-12, atol=1e-12, err_msg=f"{name} {meth}")


@pytest.mark.gpu
def test_generate_dataset_surface():
  from myriad_b200.config import Config, HParams, OptimizerType
  from myriad_b200.systems import SystemType
  from myriad_b200.utils import generate_dataset, get_state_trajectory_and_cost_batch
  import torch
  hp = HParams(system=SystemType.CARTPOLE, optimizer=OptimizerType.SHOOTING, intervals=1, controls_per_interval=25, train_size=20)
  np.random.seed(hp.seed)
  ds = generate_dataset(hp, Config(verbose=False, plot=False))
  total = hp.train_size + hp.val_size + hp.test_size
  assert ds.shape == (total, hp.num_steps + 1, 5)
  b = np.asarray(hp.system().bounds)
  assert (ds[..., 4] >= b[4, 0]).all() and (ds[..., 4] <= b[4, 1]).all()
  # the states ARE the rollouts of the controls (clipped to the state bounds like the reference does, utils.py:429)
  xs, _ = get_state_trajectory_and_cost_batch(hp, hp.system(), torch.as_tensor(ds[:, 0, :4]).cuda().contiguous(),
                                              torch.as_tensor(ds[..., 4:]).cuda().contiguous())
  np.testing.assert_allclose(ds[..., :4], np.clip(xs.cpu().numpy(), b[:4, 0], b[:4, 1]), rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "exgd_simplecase.npz")), reason="exgd fixture not generated")
def test_extragradient_matches_reference_iterates():
  """NLPSolverType.EXTRAGRADIENT: 200 steps of the reference's solver (extra_gradient.py:21-33) on its own test problem
  (tests/tests.py:253-265) reproduce the reference's iterates."""
  from myriad_b200.config import Config, HParams, IntegrationMethod, NLPSolverType, OptimizerType
  from myriad_b200.systems import SystemType
  from myriad_b200.trajectory_optimizers import get_optimizer
  fx = dict(np.load(os.path.join(GOLDEN, "exgd_simplecase.npz")))
  hp = HParams(system=SystemType.SIMPLECASE, optimizer=OptimizerType.SHOOTING, nlpsolver=NLPSolverType.EXTRAGRADIENT,
               integration_method=IntegrationMethod.HEUN, intervals=50, controls_per_interval=1, max_iter=20)
  assert hp.max_iter == 200
  opt = get_optimizer(hp, Config(verbose=False, plot=False), hp.system())
  np.testing.assert_allclose(opt.guess, fx["guess"], rtol=1e-13, atol=1e-15)
  res = opt.solve()
  np.testing.assert_allclose(res["xs_and_us"], fx["z"], rtol=1e-9, atol=1e-10)
  np.testing.assert_allclose(res["lambda"], fx["lam"], rtol=1e-9, atol=1e-10)
  np.testing.assert_allclose(res["cost"], float(fx["cost"]), rtol=1e-10)
