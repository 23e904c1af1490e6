"""Pins the CPU oracle (oracle/) against fixtures produced by the UNMODIFIED reference code
(oracle/make_golden.py ran /root/reference under oracle/refshim; fixtures in tests/golden/)."""
import numpy as np
import pytest

from oracle import nlp
from oracle.systems import make_system
from oracle.transcription import (METHODS, get_defect, get_state_trajectory_and_cost, integrate,
                                  make_transcription)
from tests.cases import CASES, load

RTOL = 1e-12


def _tr(case):
  sysname, opt, quad, meth, intervals, cpi = CASES[case]
  system = make_system(sysname)
  return system, make_transcription(system, opt, intervals, cpi, meth, quad)


def _close(a, b, rtol=RTOL, atol=1e-13):
  a, b = np.asarray(a), np.asarray(b)
  fin = np.isfinite(b)
  assert np.array_equal(np.isfinite(a), fin)
  np.testing.assert_allclose(a[fin], b[fin], rtol=rtol, atol=atol)


def test_integrator_known_answer():
  """tests/tests.py:19-43 of the reference: RK4 on y'=y reaches e to 6 decimals; all four
  methods match the reference's integrate() step for step."""
  import os
  from tests.cases import GOLDEN
  kat = dict(np.load(os.path.join(GOLDEN, "integrator_kat.npz")))
  N = 100
  t = np.linspace(0., 1., N)
  h = t[1]
  for meth in METHODS:
    _, states = integrate(lambda s, c, tt: s, np.array([1.]), t[:, None], h, N - 1, t, meth)
    _close(states, kat[meth], rtol=1e-14)
  _, states = integrate(lambda s, c, tt: s, np.array([1.]), t[:, None], h, N - 1, t, "RK4")
  np.testing.assert_almost_equal(states[-1, 0], np.e, decimal=6)


@pytest.mark.parametrize("case", sorted(CASES))
def test_transcription_matches_reference(case):
  fx = load(case)
  system, tr = _tr(case)
  _close(tr.guess, fx["guess"])
  assert np.array_equal(tr.bounds, fx["bounds"])
  assert tr.nvars == fx["guess"].shape[0] and tr.ncon == fx["con_z"].shape[0]
  _close(tr.objective(fx["guess"]), fx["obj_guess"])
  _close(tr.constraints(fx["guess"]), fx["con_guess"])
  z = fx["z"]
  _close(tr.objective(z), fx["obj_z"])
  _close(tr.constraints(z), fx["con_z"])
  _close(nlp.objective_grad(tr, z), fx["grad_z"], rtol=1e-11, atol=1e-12)
  _close(nlp.constraints_jac(tr, z), fx["jac_z"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("case", sorted(CASES))
def test_post_solve_rollout_matches_reference(case):
  fx = load(case)
  if "rollout_states" not in fx:
    pytest.skip("reference rollout not defined for this combination")
  sysname, opt, quad, meth, intervals, cpi = CASES[case]
  system, tr = _tr(case)
  if opt == "COLLOCATION":
    cpi = 1  # HParams.__post_init__ (myriad/config.py:98-99)
  _, u = tr.unravel(fx["z"])
  system = getattr(system, "true_system", system)  # a NodeSystem's verification rollout integrates the TRUE dynamics
  xs, c = get_state_trajectory_and_cost(system, intervals, cpi, meth, system.x_0, u)
  _close(xs, fx["rollout_states"], rtol=1e-11)
  _close(c, fx["rollout_cost"], rtol=1e-11)
  if "rollout_defect" in fx:
    _close(get_defect(system, xs), fx["rollout_defect"], rtol=1e-11, atol=1e-12)


def test_hessian_symmetric_and_consistent():
  """torch.func Hessian of the Lagrangian agrees with a finite difference of the complex-step gradient."""
  system, tr = _tr("s_cartpole_trap_10")
  fx = load("s_cartpole_trap_10")
  z = fx["z"]
  rng = np.random.default_rng(0)
  lam = rng.standard_normal(tr.ncon)
  H = nlp.lagrangian_hessian(tr, z, lam)
  np.testing.assert_allclose(H, H.T, atol=1e-12)

  def grad_lag(zz):
    return nlp.objective_grad(tr, zz) + nlp.constraints_jac(tr, zz).T @ lam

  v = rng.standard_normal(z.shape)
  eps = 1e-6
  fd = (grad_lag(z + eps * v) - grad_lag(z - eps * v)) / (2 * eps)
  np.testing.assert_allclose(H @ v, fd, rtol=1e-6, atol=1e-6)
