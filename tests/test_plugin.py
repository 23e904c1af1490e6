"""register_system (myriad_b200/plugin.py): the reference's system plugin point (subclass FiniteHorizonControlSystem + add
a SystemType member, myriad/systems/base.py:11-111, systems/__init__.py:29-50) for the B200 engine.  A system given as
sympy expressions is generated, compiled into its own shared library (nvcc, one translation unit: ~2-3 CPU minutes on
the first run, cached afterwards under build/user_systems/), registered through the C ABI and then used like a built-in
one.  Checked here through the HOST twin against the oracle's transcription of the same (lambdified) formulas; the GPU
variant runs the CUDA kernels of the plugin library."""
import ctypes as C
import shutil

import numpy as np
import pytest

from tests.test_parity_round2 import BACKENDS

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not __import__("os").path.exists("/usr/local/cuda/bin/nvcc"),
                                reason="nvcc is needed to build a plugin system")


def _register():
  from myriad_b200.plugin import register_system
  # damped double integrator with a quartic state penalty: x0' = x1, x1' = u - k x0 - c x1^3; cost u^2 + w x0^4
  return register_system(
    "DAMPEDINTEGRATOR", n=2, m=1, params=[("k", 0.5), ("c", 0.1), ("w", 2.0)],
    f=lambda x, u, p: [x[1], u[0] - p["k"] * x[0] - p["c"] * x[1] ** 3],
    g=lambda x, u, t, p: u[0] ** 2 + p["w"] * x[0] ** 4,
    x_0=[0.0, 0.0], x_T=[1.0, 0.0], T=2.0, bounds=[[-3.0, 3.0], [-3.0, 3.0], [-4.0, 4.0]], verbose=False)


def _oracle_system(k=0.5, c=0.1, w=2.0):
  from oracle.systems import OracleSystem, _stack

  class Damped(OracleSystem):
    def __init__(self):
      super().__init__("DAMPEDINTEGRATOR", np.zeros(2), np.array([1.0, 0.0]), 2.0,
                       np.array([[-3.0, 3.0], [-3.0, 3.0], [-4.0, 4.0]]), False, dict(k=k, c=c, w=w))

    def dynamics(self, x, u):
      return _stack([x[..., 1], u[..., 0] - k * x[..., 0] - c * x[..., 1] ** 3], x)

    def cost(self, x, u, t):
      return u[..., 0] ** 2 + w * x[..., 0] ** 4

  return Damped()


@pytest.mark.parametrize("be", BACKENDS)
def test_registered_system_matches_oracle(be):
  from myriad_b200 import problems as PR
  from oracle import nlp
  from oracle.transcription import make_transcription
  Sys = _register()
  for kw in ({}, {"k": 1.5, "w": 0.5}):
    system = Sys(**kw)
    osys = _oracle_system(**{**dict(k=0.5, c=0.1, w=2.0), **kw})
    for optid, oname, quad, meth, N, cpi in ((PR.TRAPEZOIDAL, "COLLOCATION", "TRAPEZOIDAL", "HEUN", 12, 1),
                                             (PR.HERMITE_SIMPSON, "COLLOCATION", "HERMITE_SIMPSON", "RK4", 6, 1),
                                             (PR.SHOOTING, "SHOOTING", "TRAPEZOIDAL", "HEUN", 3, 4)):
      tr = PR.Transcription(system, optid, meth, N, cpi)
      otr = make_transcription(osys, oname, N, cpi, meth, quad)
      assert tr.nvars == otr.guess.shape[0]
      rng = np.random.default_rng(0)
      z = otr.guess + 0.1 * rng.standard_normal(otr.guess.shape)
      f, c = be.eval(tr, z[None])
      np.testing.assert_allclose(f[0], float(otr.objective(z)), rtol=1e-12, atol=1e-14)
      np.testing.assert_allclose(c[0], otr.constraints(z), rtol=1e-12, atol=1e-13)
      out = be.solve(tr, otr.guess[None], otr.bounds[None, :, 0], otr.bounds[None, :, 1])
      assert int(out["status"][0]) == 0 and float(out["cinf"][0]) <= 1e-8
      ref = nlp.solve(otr, "SLSQP", ftol=1e-12, max_iter=2000)
      assert abs(float(out["obj"][0]) - ref["cost"]) <= 1e-6 * max(1.0, abs(ref["cost"])), (kw, oname, float(out["obj"][0]), ref["cost"])
      x, u = tr.unravel(out["z"][0])
      assert np.abs(u - ref["u"]).max() <= 1e-3 * 8.0


def test_registered_system_is_listed_and_rejects_bad_parameters():
  from myriad_b200 import _lib as ML
  Sys = _register()
  assert ML.SYSTEM_IDS["DAMPEDINTEGRATOR"] >= ML.USER_BASE
  with pytest.raises(TypeError):
    Sys(mass=3.0)
  s = Sys()
  assert s.state_size == 2 and s.control_size == 1 and s.device_name == "DAMPEDINTEGRATOR" and list(s.params) == [0.5, 0.1, 2.0]
