"""CPU oracle: the reference's control systems, restated.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product (myriad_b200/) never
imports it.  Parity status: pinned against the reference's own Python code executed under
oracle/refshim (see oracle/make_golden.py -> tests/golden/*.npz); solver arithmetic of IPOPT
itself is un-vendored, hence "solver-output parity unpinned vs IPOPT" (DESIGN.md).

Every function broadcasts over leading batch dimensions: ``x[..., n]``, ``u[..., m]`` and
works for NumPy (real or complex -- used for complex-step derivatives) and torch tensors
(used for torch.func Hessians).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

try:  # torch is optional for the oracle (only second derivatives need it)
  import torch
except Exception:  # pragma: no cover
  torch = None


def _xp(a):
  if torch is not None and isinstance(a, torch.Tensor):
    return torch
  return np


def _stack(xs, like):
  xp = _xp(like)
  return xp.stack(list(xs), -1) if xp is np else torch.stack(list(xs), dim=-1)


@dataclass
class OracleSystem:
  """Mirror of FiniteHorizonControlSystem (myriad/systems/base.py:11-111)."""
  name: str
  x_0: np.ndarray
  x_T: Optional[Sequence]  # entries may be None (myriad/systems/lenhart/predator_prey.py:56)
  T: float
  bounds: np.ndarray  # (n+m, 2): states first, then controls (simple_case.py:30-33)
  terminal_cost: bool = False
  params: dict = field(default_factory=dict)

  @property
  def n(self):
    return self.x_0.shape[0]

  @property
  def m(self):
    return self.bounds.shape[0] - self.n

  def dynamics(self, x, u):
    raise NotImplementedError

  def cost(self, x, u, t):
    raise NotImplementedError

  def terminal_cost_fn(self, x, u):  # base.py:101-111
    return 0.0


class SimpleCase(OracleSystem):
  """myriad/systems/lenhart/simple_case.py:25-53"""

  def __init__(self, A=1., B=1., C=4., x_0=1., T=1.):
    super().__init__("SIMPLECASE", np.array([x_0]), None, T,
                     np.array([[-np.inf, np.inf], [-np.inf, np.inf]]), False,
                     dict(A=A, B=B, C=C))

  def dynamics(self, x, u):
    C = self.params["C"]
    return _stack([-0.5 * x[..., 0] ** 2 + C * u[..., 0]], x)  # simple_case.py:48

  def cost(self, x, u, t):
    A, B = self.params["A"], self.params["B"]
    return -A * x[..., 0] + B * u[..., 0] ** 2  # simple_case.py:53


class CartPole(OracleSystem):
  """myriad/systems/classical_control/cartpole.py:50-108 (dynamics follow the CODE, :76-87,
  not the docstring at :28 -- SURVEY.md section 9-1)."""

  def __init__(self, g=9.81, m1=1., m2=.3, length=0.5):
    super().__init__("CARTPOLE", np.zeros(4), np.array([1.0, np.pi, 0., 0.]), 2.0,
                     np.array([[-2., 2.], [-2 * np.pi, 2 * np.pi], [-5., 5.], [-10., 10.],
                               [-20., 20.]]), False, dict(g=g, m1=m1, m2=m2, length=length))

  def dynamics(self, x, u):
    xp = _xp(x)
    p = self.params
    g, m1, m2, l = p["g"], p["m1"], p["m2"], p["length"]
    theta, dx, dtheta = x[..., 1], x[..., 2], x[..., 3]
    u0 = u[..., 0]
    s, c = xp.sin(theta), xp.cos(theta)
    ddx = ((l * m2 * s * dtheta ** 2 + u0 + m2 * g * c * s)
           / (m1 + m2 * (1 - c ** 2)))  # cartpole.py:79-80
    ddtheta = -((l * m2 * c * dtheta ** 2 + u0 * c + (m1 + m2) * g * s)
                / (l * m1 + l * m2 * (1 - c ** 2)))  # cartpole.py:83-85
    return _stack([dx, dtheta, ddx, ddtheta], x)

  def cost(self, x, u, t):
    return u[..., 0] ** 2  # cartpole.py:108


class VanDerPol(OracleSystem):
  """myriad/systems/miscellaneous/van_der_pol.py:29-60"""

  def __init__(self, a=1.):
    super().__init__("VANDERPOL", np.array([0., 1.]), np.zeros(2), 10.0,
                     np.array([[-4., 4.], [-4., 4.], [-0.75, 1.0]]), False, dict(a=a))

  def dynamics(self, x, u):
    a = self.params["a"]
    x0, x1 = x[..., 0], x[..., 1]
    return _stack([a * (1. - x1 ** 2) * x0 - x1 + u[..., 0], x0], x)  # van_der_pol.py:48-50

  def cost(self, x, u, t):
    return x[..., 0] ** 2 + x[..., 1] ** 2 + u[..., 0] ** 2  # van_der_pol.py:60 (x.T @ x + u**2)


class CancerTreatment(OracleSystem):
  """myriad/systems/lenhart/cancer_treatment.py:40-76"""

  def __init__(self, r=0.3, a=3., delta=0.45, x_0=0.975, T=20):
    super().__init__("CANCERTREATMENT", np.array([x_0]), None, T,
                     np.array([[1e-3, 1.], [0., 2.]]), False, dict(r=r, a=a, delta=delta))

  def dynamics(self, x, u):
    xp = _xp(x)
    r, delta = self.params["r"], self.params["delta"]
    x0 = x[..., 0]
    return _stack([r * x0 * xp.log(1 / x0) - u[..., 0] * delta * x0], x)  # cancer_treatment.py:64

  def cost(self, x, u, t):
    return self.params["a"] * x[..., 0] ** 2 + u[..., 0] ** 2  # cancer_treatment.py:76


class MouldFungicide(OracleSystem):
  """myriad/systems/lenhart/mould_fungicide.py:29-66"""

  def __init__(self, r=0.3, M=10., A=10., x_0=1.0, T=5):
    super().__init__("MOULDFUNGICIDE", np.array([x_0]), None, T, np.array([[0., 5.], [0., 5.]]), False, dict(r=r, M=M, A=A))

  def dynamics(self, x, u):
    p = self.params
    return _stack([p["r"] * (p["M"] - x[..., 0]) - u[..., 0] * x[..., 0]], x)  # mould_fungicide.py:52

  def cost(self, x, u, t):
    return self.params["A"] * x[..., 0] ** 2 + u[..., 0] ** 2  # mould_fungicide.py:66


class Bioreactor(OracleSystem):
  """myriad/systems/lenhart/bioreactor.py:36-83"""

  def __init__(self, K=2., G=1., D=1., M=1., x_0=(.5, .1), T=2.):
    super().__init__("BIOREACTOR", np.array([x_0[0]]), None, T, np.array([[0., 1.], [0., M]]), False, dict(K=K, G=G, D=D))

  def dynamics(self, x, u):
    p = self.params
    return _stack([p["G"] * u[..., 0] * x[..., 0] - p["D"] * x[..., 0] ** 2], x)  # bioreactor.py:67

  def cost(self, x, u, t):
    return -self.params["K"] * x[..., 0] + u[..., 0]  # bioreactor.py:83


class SimpleCaseWithBounds(OracleSystem):
  """myriad/systems/lenhart/simple_case_with_bounds.py:25-56"""

  def __init__(self, A=1., C=4., M_1=-1., M_2=2., x_0=1., T=1.):
    super().__init__("SIMPLECASEWITHBOUNDS", np.array([x_0]), None, T, np.array([[0., 3.], [M_1, M_2]]), False, dict(A=A, C=C))

  def dynamics(self, x, u):
    return _stack([-0.5 * x[..., 0] ** 2 + self.params["C"] * u[..., 0]], x)  # simple_case_with_bounds.py:51

  def cost(self, x, u, t):
    return -self.params["A"] * x[..., 0] + u[..., 0] ** 2  # simple_case_with_bounds.py:56


class Glucose(OracleSystem):
  """myriad/systems/lenhart/glucose.py:43-104"""

  def __init__(self, a=1., b=1., c=1., A=2., l=.5, x_0=(.75, 0.), T=.2):
    super().__init__("GLUCOSE", np.array([x_0[0], x_0[1]]), None, T, np.array([[0., 1.], [0., 1.], [0., 0.01]]), False,
                     dict(a=a, b=b, c=c, A=A, l=l))

  def dynamics(self, x, u):
    p = self.params
    return _stack([-p["a"] * x[..., 0] - p["b"] * x[..., 1], -p["c"] * x[..., 1] + u[..., 0]], x)  # glucose.py:78-81

  def cost(self, x, u, t):
    p = self.params
    return 100_000 * (p["A"] * (x[..., 0] - p["l"]) ** 2 + u[..., 0] ** 2)  # glucose.py:104


class Harvest(OracleSystem):
  """myriad/systems/lenhart/harvest.py:32-62 (time-dependent cost)"""

  def __init__(self, A=5., k=10., m=.2, M=1., x_0=.4, T=10.):
    super().__init__("HARVEST", np.array([x_0]), None, T, np.array([[-np.inf, np.inf], [0., M]]), False, dict(A=A, k=k, m=m))

  def dynamics(self, x, u):
    return _stack([-(self.params["m"] + u[..., 0]) * x[..., 0]], x)  # harvest.py:57

  def cost(self, x, u, t):
    p = self.params
    return -p["A"] * (p["k"] * t / (t + 1)) * x[..., 0] * u[..., 0] + u[..., 0] ** 2  # harvest.py:62


class TimberHarvest(OracleSystem):
  """myriad/systems/lenhart/timber_harvest.py:41-85 (time-dependent cost)"""

  def __init__(self, r=0., k=1., x_0=100., T=5.):
    super().__init__("TIMBERHARVEST", np.array([x_0]), None, T, np.array([[0., 20_000.], [0., 1.]]), False, dict(r=r, k=k))

  def dynamics(self, x, u):
    return _stack([self.params["k"] * x[..., 0] * u[..., 0]], x)  # timber_harvest.py:64

  def cost(self, x, u, t):
    xp = _xp(x)
    e = xp.exp(-self.params["r"] * (t if xp is np else torch.as_tensor(t, dtype=x.dtype)))
    return -e * x[..., 0] * (1 - u[..., 0])  # timber_harvest.py:85


class SEIR(OracleSystem):
  """myriad/systems/miscellaneous/seir.py:47-95"""

  def __init__(self, A=0.1, b=0.525, d=0.5, c=0.0001, e=0.5, g=0.1, a=0.2):
    super().__init__("SEIR", np.array([1000.0, 100.0, 50.0, 1165.0]), None, 20,
                     np.array([[0., 2000.], [0., 250.], [0., 250.], [0., 3000.], [0., 1.]]), False,
                     dict(A=A, b=b, d=d, c=c, e=e, g=g, a=a))

  def dynamics(self, x, u):
    p = self.params
    S_, E_, I_, N_ = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    u0 = u[..., 0]
    return _stack([p["b"] * N_ - p["d"] * S_ - p["c"] * S_ * I_ - u0 * S_,       # seir.py:86
                   p["c"] * S_ * I_ - (p["e"] + p["d"]) * E_,                   # :87
                   p["e"] * E_ - (p["g"] + p["a"] + p["d"]) * I_,               # :88
                   (p["b"] - p["d"]) * N_ - p["a"] * I_], x)                    # :89

  def cost(self, x, u, t):
    return self.params["A"] * x[..., 2] + u[..., 0] ** 2  # seir.py:95


class EpidemicSEIRN(SEIR):
  """myriad/systems/lenhart/epidemic_seirn.py:43-95: same vector field, unbounded states, u in [0, 0.9]"""

  def __init__(self, A=.1, b=.525, d=.5, c=.0001, e=.5, g=.1, a=.2, x_0=(1000., 100., 50., 15.), T=20.):
    OracleSystem.__init__(self, "EPIDEMICSEIRN", np.array([x_0[0], x_0[1], x_0[2], float(np.sum(x_0))]), None, T,
                          np.array([[-np.inf, np.inf]] * 4 + [[0., 0.9]]), False, dict(A=A, b=b, d=d, c=c, e=e, g=g, a=a))


class HIVTreatment(OracleSystem):
  """myriad/systems/lenhart/hiv_treatment.py:33-111"""

  def __init__(self, s=10., m_1=.02, m_2=.5, m_3=4.4, r=.03, T_max=1500., k=.000024, N=300., x_0=(800., .04, 1.5), A=.05, T=20.):
    super().__init__("HIVTREATMENT", np.array([x_0[0], x_0[1], x_0[2]]), None, T,
                     np.array([[0., 1600.], [0., 100.], [0., 100.], [0., 1.]]), False,
                     dict(s=s, m_1=m_1, m_2=m_2, m_3=m_3, r=r, T_max=T_max, k=k, N=N, A=A))

  def dynamics(self, x, u):
    p = self.params
    x0, x1, x2 = x[..., 0], x[..., 1], x[..., 2]
    u0 = u[..., 0]
    return _stack([p["s"] / (1 + x2) - p["m_1"] * x0 + p["r"] * x0 * (1 - (x0 + x1) / p["T_max"]) - u0 * p["k"] * x0 * x2,  # :79
                   u0 * p["k"] * x0 * x2 - p["m_2"] * x1,                       # :80
                   p["N"] * p["m_2"] * x1 - p["m_3"] * x2], x)                  # :81

  def cost(self, x, u, t):
    return -self.params["A"] * x[..., 0] + (1 - u[..., 0]) ** 2  # hiv_treatment.py:111


class Bacteria(OracleSystem):
  """myriad/systems/lenhart/bacteria.py:34-86"""

  def __init__(self, r=1., A=1., B=12., C=1., x_0=1.):
    super().__init__("BACTERIA", np.array([x_0]), None, 1, np.array([[0., 10.], [0., 2.]]), True, dict(r=r, A=A, B=B, C=C))

  def dynamics(self, x, u):
    xp = _xp(x)
    p = self.params
    x0, u0 = x[..., 0], u[..., 0]
    return _stack([p["r"] * x0 + p["A"] * u0 * x0 - p["B"] * u0 ** 2 * xp.exp(-x0)], x)  # bacteria.py:61

  def cost(self, x, u, t):
    return u[..., 0] ** 2  # bacteria.py:77

  def terminal_cost_fn(self, x, u):
    return -self.params["C"] * x[..., 0]  # bacteria.py:86


class Tumour(OracleSystem):
  """myriad/systems/miscellaneous/tumour.py:44-108"""

  def __init__(self, xi=0.084, b=5.85, d=0.00873, G=0.15, mu=0.02):
    p_ = ((b - mu) / d) ** (3 / 2)
    super().__init__("TUMOUR", np.array([p_ / 2, p_ / 4, 0.0]), None, 1.2,
                     np.array([[0., p_], [0., p_], [0., 15.], [0., 75.]]), True, dict(xi=xi, b=b, d=d, G=G, mu=mu))

  def dynamics(self, x, u):
    xp = _xp(x)
    p = self.params
    pp, q, u0 = x[..., 0], x[..., 1], u[..., 0]
    return _stack([-p["xi"] * pp * xp.log(pp / q),                                             # tumour.py:79
                   q * (p["b"] - (p["mu"] + p["d"] * pp ** (2 / 3) + p["G"] * u0)),            # :80
                   u0], x)                                                            # :81

  def cost(self, x, u, t):
    return 0 * x[..., 0]  # tumour.py:98

  def terminal_cost_fn(self, x, u):
    return x[..., 0]  # tumour.py:106-108


class PredatorPrey(OracleSystem):
  """myriad/systems/lenhart/predator_prey.py:47-122"""

  def __init__(self, d_1=.1, d_2=.1, A=1., B=5., M=1., x_0=(10., 1., 0.), T=10.):
    super().__init__("PREDATORPREY", np.array([x_0[0], x_0[1], x_0[2]]), [None, None, B], T,
                     np.array([[0., 11.], [0., 11.], [0., 5.], [0., M]]), True, dict(d_1=d_1, d_2=d_2, A=A))

  def dynamics(self, x, u):
    p = self.params
    x0, x1, u0 = x[..., 0], x[..., 1], u[..., 0]
    return _stack([(1 - x1) * x0 - p["d_1"] * x0 * u0, (x0 - 1) * x1 - p["d_2"] * x1 * u0, u0], x)  # :86-90

  def cost(self, x, u, t):
    return self.params["A"] * 0.5 * u[..., 0] ** 2  # predator_prey.py:114

  def terminal_cost_fn(self, x, u):
    return x[..., 0]  # predator_prey.py:122


class BearPopulations(OracleSystem):
  """myriad/systems/lenhart/bear_populations.py:36-110"""

  def __init__(self, r=.1, K=.75, m_p=.5, m_f=.5, c_p=10_000, c_f=10, x_0=(.4, .2, 0.), T=25):
    super().__init__("BEARPOPULATIONS", np.array([x_0[0], x_0[1], x_0[2]]), None, T,
                     np.array([[0., 2.], [0., 2.], [0., 2.], [0., .2], [0., .2]]), False,
                     dict(r=r, K=K, m_p=m_p, m_f=m_f, c_p=c_p, c_f=c_f))

  def dynamics(self, x, u):
    p = self.params
    r, K, m_p, m_f = p["r"], p["K"], p["m_p"], p["m_f"]
    k, k2 = r / K, r / K ** 2
    x0, x1 = x[..., 0], x[..., 1]
    u0, u1 = u[..., 0], u[..., 1]
    return _stack([r * x0 - k * x0 ** 2 + k * m_f * (1 - x0 / K) * x1 ** 2 - u0 * x0,                          # :78
                   r * x1 - k * x1 ** 2 + k * m_p * (1 - x1 / K) * x0 ** 2 - u1 * x1,                          # :79
                   k * (1 - m_p) * x0 ** 2 + k * (1 - m_f) * x1 ** 2 + k2 * m_f * x0 * x1 ** 2 + k2 * m_p * (x0 ** 2) * x1], x)  # :80-81

  def cost(self, x, u, t):
    p = self.params
    return x[..., 2] + p["c_p"] * u[..., 0] ** 2 + p["c_f"] * u[..., 1] ** 2  # bear_populations.py:110


class NodeSystem(OracleSystem):
  """NODE-dynamics wrapper: myriad/systems/neural_ode/node_system.py:14-42 with the MLP of
  myriad/neural_ode/create_node.py:110-117 (hk.Linear = x @ w + b, sigmoid between layers).
  ``weights`` is a list of (w:(in,out), b:(out,)) in layer order (haiku keys linear, linear_1, ...)."""

  def __init__(self, true_system: OracleSystem, weights):
    super().__init__("NODE_" + true_system.name, true_system.x_0, true_system.x_T, true_system.T,
                     true_system.bounds, true_system.terminal_cost, {})
    self.true_system = true_system
    self.weights = [(np.asarray(w, dtype=np.float64), np.asarray(b, dtype=np.float64)) for w, b in weights]

  def dynamics(self, x, u):
    xp = _xp(x)
    if xp is np:
      h = np.concatenate([x, u], axis=-1)  # jnp.append(x_t, u_t), node_system.py:37
      for i, (w, b) in enumerate(self.weights):
        h = h @ w + b
        if i + 1 < len(self.weights):
          h = 1.0 / (1.0 + np.exp(-h))
      return h
    h = torch.cat([x, u], dim=-1)
    for i, (w, b) in enumerate(self.weights):
      h = h @ torch.as_tensor(w) + torch.as_tensor(b)
      if i + 1 < len(self.weights):
        h = torch.sigmoid(h)
    return h

  def cost(self, x, u, t):
    return self.true_system.cost(x, u, t)  # node_system.py:41-42


def haiku_style_mlp_weights(n_in: int, hidden: Sequence[int], n_out: int, seed: int = 42):
  """Synthetic MLP weights with haiku's default Linear init (truncated normal, stddev 1/sqrt(fan_in),
  zero bias), drawn with numpy's PCG64 so they are reproducible without jax."""
  rng = np.random.Generator(np.random.PCG64(seed))
  sizes = [n_in] + list(hidden) + [n_out]
  out = []
  for fi, fo in zip(sizes[:-1], sizes[1:]):
    std = 1.0 / math.sqrt(fi)
    w = rng.standard_normal((fi, fo))
    w = np.clip(w, -2.0, 2.0) * std
    out.append((w, np.zeros(fo)))
  return out


def _clip(v, lo, hi):
  """jnp.clip, also for complex-step inputs (derivative 1 strictly inside, 0 outside) and torch tensors"""
  if torch is not None and isinstance(v, torch.Tensor):
    return torch.clamp(v, lo, hi)
  v = np.asarray(v)
  if np.iscomplexobj(v):
    return np.where(v.real < lo, lo + 0j, np.where(v.real > hi, hi + 0j, v))
  return np.clip(v, lo, hi)


def _angle_normalize(x):
  """pendulum.py:17-18: ((x + pi) % (2 pi)) - pi with jnp.remainder semantics"""
  if torch is not None and isinstance(x, torch.Tensor):
    return torch.remainder(x + math.pi, 2 * math.pi) - math.pi
  x = np.asarray(x)
  re = x.real if np.iscomplexobj(x) else x
  return x + math.pi - np.floor((re + math.pi) / (2 * math.pi)) * (2 * math.pi) - math.pi


class RocketLanding(OracleSystem):
  """myriad/systems/miscellaneous/rocket_landing.py:54-124"""

  def __init__(self, g=9.8, m=100_000, length=50, width=10):
    deg = 0.01745329
    super().__init__("ROCKETLANDING", np.array([0., 0., 1000., -80., -np.pi / 2., 0.]), np.zeros(6), 16.,
                     np.array([[-250., 150.], [-250., 150.], [0., 1000.], [-250., 150.], [-2 * np.pi, 2 * np.pi], [-250., 150.],
                               [0.4, 1.], [-20 * deg, 20 * deg]]), False,
                     dict(g=g, m=m, length=length, max_thrust=1 * 2210 * 1000))

  def dynamics(self, x, u):
    xp = _xp(x)
    p = self.params
    theta = x[..., 4]
    thrust, ang = u[..., 0], u[..., 1]
    Fx = p["max_thrust"] * thrust * xp.sin(ang + theta)           # rocket_landing.py:107
    Fy = p["max_thrust"] * thrust * xp.cos(ang + theta)           # :112
    Tq = -p["length"] / 2 * p["max_thrust"] * thrust * xp.sin(ang)  # :117
    I = 1 / 12 * p["m"] * p["length"] ** 2                        # :62
    return _stack([x[..., 1], Fx / p["m"], x[..., 3], Fy / p["m"] - p["g"], x[..., 5], Tq / I], x)

  def cost(self, x, u, t):
    return u[..., 0] ** 2 + u[..., 1] ** 2 + 2 * x[..., 5] ** 2   # :124


class Pendulum(OracleSystem):
  """myriad/systems/classical_control/pendulum.py:50-119 (clipped torque / speed, normalised angle)"""

  def __init__(self, g=10., m=1., length=1.):
    super().__init__("PENDULUM", np.array([0., 0.]), np.array([np.pi, 0.]), 15.,
                     np.array([[-np.pi, np.pi], [-8., 8.], [-2., 2.]]), False,
                     dict(g=g, m=m, length=length, max_speed=8., max_torque=2., ctrl_penalty=0.001))

  def dynamics(self, x, u):
    xp = _xp(x)
    p = self.params
    uu = _clip(u[..., 0], -p["max_torque"], p["max_torque"])      # pendulum.py:97
    theta = _angle_normalize(x[..., 0])                            # :100
    dth = _clip(x[..., 1], -p["max_speed"], p["max_speed"])       # :101
    ddth = (-3. * p["g"] / (2. * p["length"]) * xp.sin(theta) + 3. * uu / (p["m"] * p["length"] ** 2)) * 0.05  # :104-105
    return _stack([dth, ddth], x)

  def cost(self, x, u, t):
    p = self.params
    return _angle_normalize(x[..., 0]) ** 2 + 0.1 * x[..., 1] ** 2 + p["ctrl_penalty"] * u[..., 0] ** 2  # :119


class MountainCar(OracleSystem):
  """myriad/systems/classical_control/mountain_car.py:53-107; hill_function(x) = x^2 / 2 (:11-13)"""

  def __init__(self, power=0.0015, gravity=0.0025):
    super().__init__("MOUNTAINCAR", np.array([-0.1, 0.]), np.array([0.45, 0.]), 300.,
                     np.array([[-1.2, 0.6], [-0.07, 0.07], [-1., 1.]]), False, dict(power=power, gravity=gravity))

  def dynamics(self, x, u):
    p = self.params
    force = _clip(u[..., 0], -1.0, 1.0)                            # mountain_car.py:88
    return _stack([x[..., 1], force * p["power"] - p["gravity"] * x[..., 0]], x)  # :90-91

  def cost(self, x, u, t):
    return 10. * u[..., 0] ** 2                                    # :103


SYSTEMS = {
  "SIMPLECASE": SimpleCase,
  "CARTPOLE": CartPole,
  "VANDERPOL": VanDerPol,
  "CANCERTREATMENT": CancerTreatment,
  "MOULDFUNGICIDE": MouldFungicide,
  "BIOREACTOR": Bioreactor,
  "SIMPLECASEWITHBOUNDS": SimpleCaseWithBounds,
  "GLUCOSE": Glucose,
  "HARVEST": Harvest,
  "TIMBERHARVEST": TimberHarvest,
  "SEIR": SEIR,
  "EPIDEMICSEIRN": EpidemicSEIRN,
  "HIVTREATMENT": HIVTreatment,
  "BACTERIA": Bacteria,
  "TUMOUR": Tumour,
  "PREDATORPREY": PredatorPrey,
  "BEARPOPULATIONS": BearPopulations,
  "ROCKETLANDING": RocketLanding,
  "PENDULUM": Pendulum,
  "MOUNTAINCAR": MountainCar,
}


def golden_node_weights(name: str):
  """[(w, b), ...] of the committed NODE fixture weights for system ``NODE_<true system>`` (tests/golden/node_*.npz,
  made by tools/fit_node.py), haiku layer order linear, linear_1, ... (create_node.py:124-131)."""
  import os
  files = {"NODE_CARTPOLE": "node_cartpole_64x64x64.npz"}
  path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", files[name])
  d = dict(np.load(path))
  out, i = [], 0
  while ("linear" if i == 0 else f"linear_{i}") + "/w" in d:
    k = "linear" if i == 0 else f"linear_{i}"
    out.append((d[k + "/w"], d[k + "/b"]))
    i += 1
  return out


def make_system(name: str, **params) -> OracleSystem:
  if name.startswith("NODE_"):
    return NodeSystem(SYSTEMS[name[5:]](**params), golden_node_weights(name))
  return SYSTEMS[name](**params)
