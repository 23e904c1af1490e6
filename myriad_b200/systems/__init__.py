"""SystemType plugin surface (mirror of myriad/systems/__init__.py:29-53): enum members are callable
constructors with the reference's constructor arguments and defaults."""
from __future__ import annotations

from enum import Enum

import numpy as np

from .base import FiniteHorizonControlSystem


class SimpleCase(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/simple_case.py:25-53"""

  def __init__(self, A=1., B=1., C=4., x_0=1., T=1.):
    super().__init__(x_0=np.array([x_0], dtype=np.float64), x_T=None, T=T,
                     bounds=np.array([[-np.inf, np.inf], [-np.inf, np.inf]]), terminal_cost=False, discrete=False,
                     device_name="SIMPLECASE", params=[A, B, C])
    self.A, self.B, self.C = A, B, C


class CartPole(FiniteHorizonControlSystem):
  """myriad/systems/classical_control/cartpole.py:50-73"""

  def __init__(self, g: float = 9.81, m1: float = 1., m2: float = .3, length: float = 0.5):
    self.m1, self.m2, self.length, self.g = m1, m2, length, g
    self.u_max, self.d_max, self.d = 20, 2.0, 1.0
    super().__init__(x_0=np.array([0., 0., 0., 0.]), x_T=np.array([self.d, np.pi, 0., 0.]), T=2.0,
                     bounds=np.array([[-self.d_max, self.d_max], [-2 * np.pi, 2 * np.pi], [-5., 5.], [-10., 10.],
                                      [-self.u_max, self.u_max]]),
                     terminal_cost=False, device_name="CARTPOLE", params=[g, m1, m2, length])


class VanDerPol(FiniteHorizonControlSystem):
  """myriad/systems/miscellaneous/van_der_pol.py:29-44"""

  def __init__(self, a=1.):
    self.a = a
    super().__init__(x_0=np.array([0., 1.]), x_T=np.zeros(2), T=10.0,
                     bounds=np.array([[-4., 4.], [-4., 4.], [-0.75, 1.0]]), terminal_cost=False,
                     device_name="VANDERPOL", params=[a])


class CancerTreatment(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/cancer_treatment.py:40-60"""

  def __init__(self, r=0.3, a=3., delta=0.45, x_0=0.975, T=20):
    super().__init__(x_0=np.array([x_0], dtype=np.float64), x_T=None, T=T, bounds=np.array([[1e-3, 1.], [0., 2.]]),
                     terminal_cost=False, discrete=False, device_name="CANCERTREATMENT", params=[r, a, delta])
    self.r, self.a, self.delta = r, a, delta


class MouldFungicide(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/mould_fungicide.py:29-48"""

  def __init__(self, r=0.3, M=10., A=10., x_0=1.0, T=5):
    super().__init__(x_0=np.array([x_0], dtype=np.float64), x_T=None, T=T, bounds=np.array([[0., 5.], [0., 5.]]),
                     terminal_cost=False, discrete=False, device_name="MOULDFUNGICIDE", params=[r, M, A])


class SimpleCaseWithBounds(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/simple_case_with_bounds.py:25-47"""

  def __init__(self, A=1., C=4., M_1=-1., M_2=2., x_0=1., T=1.):
    super().__init__(x_0=np.array([x_0], dtype=np.float64), x_T=None, T=T, bounds=np.array([[0., 3.], [M_1, M_2]]),
                     terminal_cost=False, discrete=False, device_name="SIMPLECASEWITHBOUNDS", params=[A, C])


class Bioreactor(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/bioreactor.py:36-83"""

  def __init__(self, K=2., G=1., D=1., M=1., x_0=(.5, .1), T=2.):
    super().__init__(x_0=np.array([x_0[0]], dtype=np.float64), x_T=None, T=T, bounds=np.array([[0., 1.], [0., M]]),
                     terminal_cost=False, discrete=False, device_name="BIOREACTOR", params=[K, G, D])


class Glucose(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/glucose.py:43-104"""

  def __init__(self, a=1., b=1., c=1., A=2., l=.5, x_0=(.75, 0.), T=.2):
    super().__init__(x_0=np.array([x_0[0], x_0[1]], dtype=np.float64), x_T=None, T=T,
                     bounds=np.array([[0., 1.], [0., 1.], [0., 0.01]]), terminal_cost=False, discrete=False,
                     device_name="GLUCOSE", params=[a, b, c, A, l])


class Harvest(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/harvest.py:32-62 (time-dependent running cost)"""

  def __init__(self, A=5., k=10., m=.2, M=1., x_0=.4, T=10.):
    super().__init__(x_0=np.array([x_0], dtype=np.float64), x_T=None, T=T, bounds=np.array([[-np.inf, np.inf], [0., M]]),
                     terminal_cost=False, discrete=False, device_name="HARVEST", params=[A, k, m])


class TimberHarvest(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/timber_harvest.py:41-85 (time-dependent running cost)"""

  def __init__(self, r=0., k=1., x_0=100., T=5.):
    super().__init__(x_0=np.array([x_0], dtype=np.float64), x_T=None, T=T, bounds=np.array([[0., 20_000.], [0., 1.]]),
                     terminal_cost=False, discrete=False, device_name="TIMBERHARVEST", params=[r, k])


class SEIR(FiniteHorizonControlSystem):
  """myriad/systems/miscellaneous/seir.py:47-95 (parameters are fixed in the reference's constructor)"""

  def __init__(self):
    self.b, self.d, self.c, self.e, self.g, self.a, self.A = 0.525, 0.5, 0.0001, 0.5, 0.1, 0.2, 0.1
    super().__init__(x_0=np.array([1000.0, 100.0, 50.0, 1000.0 + 100.0 + 50.0 + 15.0]), x_T=None, T=20,
                     bounds=np.array([[0., 2000.], [0., 250.], [0., 250.], [0., 3000.], [0., 1.]]), terminal_cost=False,
                     device_name="SEIR", params=[self.A, self.b, self.d, self.c, self.e, self.g, self.a])


class EpidemicSEIRN(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/epidemic_seirn.py:43-95"""

  def __init__(self, A=.1, b=.525, d=.5, c=.0001, e=.5, g=.1, a=.2, x_0=(1000., 100., 50., 15.), T=20.):
    inf = np.inf
    super().__init__(x_0=np.array([x_0[0], x_0[1], x_0[2], float(np.sum(np.asarray(x_0)))]), x_T=None, T=T,
                     bounds=np.array([[-inf, inf], [-inf, inf], [-inf, inf], [-inf, inf], [0., 0.9]]), terminal_cost=False,
                     discrete=False, device_name="EPIDEMICSEIRN", params=[A, b, d, c, e, g, a])


class HIVTreatment(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/hiv_treatment.py:33-111"""

  def __init__(self, s=10., m_1=.02, m_2=.5, m_3=4.4, r=.03, T_max=1500., k=.000024, N=300., x_0=(800., .04, 1.5), A=.05, T=20.):
    super().__init__(x_0=np.array([x_0[0], x_0[1], x_0[2]], dtype=np.float64), x_T=None, T=T,
                     bounds=np.array([[0., 1600.], [0., 100.], [0., 100.], [0., 1.]]), terminal_cost=False, discrete=False,
                     device_name="HIVTREATMENT", params=[s, m_1, m_2, m_3, r, T_max, k, N, A])


class Bacteria(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/bacteria.py:34-86 (terminal cost -C x(T))"""

  def __init__(self, r=1., A=1., B=12., C=1., x_0=1.):
    super().__init__(x_0=np.array([x_0], dtype=np.float64), x_T=None, T=1, bounds=np.array([[0., 10.], [0., 2.]]),
                     terminal_cost=True, discrete=False, device_name="BACTERIA", params=[r, A, B, C])
    self.r, self.A, self.B, self.C = r, A, B, C

  def terminal_cost_fn(self, x_T, u_T, T=None):
    return -self.C * np.squeeze(x_T)


class Tumour(FiniteHorizonControlSystem):
  """myriad/systems/miscellaneous/tumour.py:44-108 (no running cost; terminal cost p(T))"""

  def __init__(self, xi=0.084, b=5.85, d=0.00873, G=0.15, mu=0.02):
    p_ = q_ = ((b - mu) / d) ** (3 / 2)
    super().__init__(x_0=np.array([p_ / 2, q_ / 4, 0.0]), x_T=None, T=1.2,
                     bounds=np.array([[0., p_], [0., q_], [0., 15.], [0., 75.]]), terminal_cost=True,
                     device_name="TUMOUR", params=[xi, b, d, G, mu])

  def terminal_cost_fn(self, x_T, u_T, T=None):
    return x_T[0]


class PredatorPrey(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/predator_prey.py:47-122: only the third state has a terminal value (x_T = [None, None, B]);
  like in the reference, both collocation transcriptions need a fully specified x_T (trapezoidal.py:70-71, hermite_simpson.py:37-48
  raise on the None entries) -- use SHOOTING."""

  def __init__(self, d_1=.1, d_2=.1, A=1., B=5., guess_a=-.52, guess_b=.5, M=1., x_0=(10., 1., 0.), T=10.):
    super().__init__(x_0=np.array([x_0[0], x_0[1], x_0[2]], dtype=np.float64), x_T=[None, None, B], T=T,
                     bounds=np.array([[0., 11.], [0., 11.], [0., 5.], [0., M]]), terminal_cost=True, discrete=False,
                     device_name="PREDATORPREY", params=[d_1, d_2, A])
    self.guess_a, self.guess_b = guess_a, guess_b  # secant start values of the FBSM (predator_prey.py:75-77)

  def terminal_cost_fn(self, x_T, u_T, T=None):
    return x_T[0]


class BearPopulations(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/bear_populations.py:36-110 (two controls)"""

  def __init__(self, r=.1, K=.75, m_p=.5, m_f=.5, c_p=10_000, c_f=10, x_0=(.4, .2, 0.), T=25):
    super().__init__(x_0=np.array([x_0[0], x_0[1], x_0[2]], dtype=np.float64), x_T=None, T=T,
                     bounds=np.array([[0., 2.], [0., 2.], [0., 2.], [0., .2], [0., .2]]), terminal_cost=False, discrete=False,
                     device_name="BEARPOPULATIONS", params=[r, K, m_p, m_f, c_p, c_f])


class RocketLanding(FiniteHorizonControlSystem):
  """myriad/systems/miscellaneous/rocket_landing.py:54-124 (six states, two controls)"""

  def __init__(self, g: float = 9.8, m: float = 100_000, length: float = 50, width: float = 10) -> None:
    self.g, self.m, self.length, self.width = g, m, length, width
    self.min_thrust, self.max_thrust = 880 * 1000, 1 * 2210 * 1000
    self.I = 1 / 12 * m * length ** 2
    deg_to_rad = 0.01745329
    self.max_gimble = 20 * deg_to_rad
    self.min_gimble = -self.max_gimble
    self.min_percent_thrust, self.max_percent_thrust = 0.4, 1.
    super().__init__(x_0=np.array([0., 0., 1000., -80., -np.pi / 2., 0.]), x_T=np.array([0., 0., 0., 0., 0., 0.]), T=16.,
                     bounds=np.array([[-250., 150.], [-250., 150.], [0., 1000.], [-250., 150.], [-2 * np.pi, 2 * np.pi],
                                      [-250., 150.], [self.min_percent_thrust, self.max_percent_thrust],
                                      [self.min_gimble, self.max_gimble]]),
                     terminal_cost=False, device_name="ROCKETLANDING", params=[g, m, length, self.max_thrust])


class Pendulum(FiniteHorizonControlSystem):
  """myriad/systems/classical_control/pendulum.py:50-119.  Non-smooth (clipped torque and speed, normalised angle):
  the generated device code takes JAX's derivative choices (clip: 1 strictly inside, 0 outside; remainder: 1)."""

  def __init__(self, g: float = 10., m: float = 1., length: float = 1.):
    self.g, self.m, self.length = g, m, length
    self.max_speed, self.max_torque, self.ctrl_penalty = 8., 2., 0.001
    super().__init__(x_0=np.array([0., 0.]), x_T=np.array([np.pi, 0.]), T=15,
                     bounds=np.array([[-np.pi, np.pi], [-self.max_speed, self.max_speed], [-self.max_torque, self.max_torque]]),
                     terminal_cost=False, device_name="PENDULUM",
                     params=[g, m, length, self.max_speed, self.max_torque, self.ctrl_penalty])


class MountainCar(FiniteHorizonControlSystem):
  """myriad/systems/classical_control/mountain_car.py:53-107 (hill_function(x) = x^2 / 2, :11-13; clipped force)"""

  def __init__(self, power=0.0015, gravity=0.0025) -> None:
    self.power, self.gravity = power, gravity
    super().__init__(x_0=np.array([-0.1, 0.]), x_T=np.array([0.45, 0.]), T=300.,
                     bounds=np.array([[-1.2, 0.6], [-0.07, 0.07], [-1.0, 1.0]]), terminal_cost=False,
                     device_name="MOUNTAINCAR", params=[power, gravity])


class NodeSystem(FiniteHorizonControlSystem):
  """myriad/systems/neural_ode/node_system.py:14-42: a system whose (parametrized) dynamics is the neural-ODE MLP of
  myriad/neural_ode/create_node.py:110-117 applied to concat(x, u), while cost, bounds, horizon, start/end states and
  the post-solve verification rollout are the true system's.

  ``node`` is anything with ``.params`` -- the haiku parameter mapping {'linear': {'w','b'}, 'linear_1': ..., ...}
  (create_node.py:124-131; flat 'linear/w' keys as in an .npz are accepted too) -- e.g. myriad_b200.neural_ode.NeuralODE.
  Planning with it (get_optimizer(hp, cfg, NodeSystem(node, system)).solve()) is what the reference does through
  plan_with_node_model (myriad/utils.py:230-242)."""

  def __init__(self, node, true_system: FiniteHorizonControlSystem) -> None:
    self.node = node
    self.true_system = true_system
    super().__init__(x_0=true_system.x_0, x_T=true_system.x_T, T=true_system.T, bounds=true_system.bounds,
                     terminal_cost=true_system.terminal_cost, discrete=true_system.discrete,
                     device_name="NODE_" + true_system.device_name, params=list(true_system.params))
    self.layers = mlp_layers(node.params if hasattr(node, "params") else node)
    n_in = self.state_size + self.control_size
    sizes = [w.shape for w, _ in self.layers]
    if sizes[0][0] != n_in or sizes[-1][1] != self.state_size or any(a[1] != b[0] for a, b in zip(sizes[:-1], sizes[1:])):
      raise ValueError(f"MLP layer shapes {sizes} do not map {n_in} inputs to {self.state_size} outputs")
    self.hidden = [int(w.shape[1]) for w, _ in self.layers[:-1]]
    # flat weight vector in the order the C ABI documents: per layer w (in, out) row-major, then b
    self.theta = np.concatenate([np.concatenate([np.asarray(w, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()])
                                 for w, b in self.layers])
    self._theta_dev = {}

  def theta_device(self, device):
    import torch
    key = str(device)
    if key not in self._theta_dev:
      self._theta_dev[key] = torch.as_tensor(self.theta, dtype=torch.float64).to(device).contiguous()
    return self._theta_dev[key]


def mlp_layers(params):
  """haiku parameter mapping -> [(w (in,out), b (out,)), ...] in layer order linear, linear_1, linear_2, ..."""
  def get(name):
    if name in params:
      p = params[name]
      return np.asarray(p['w'], dtype=np.float64), np.asarray(p['b'], dtype=np.float64)
    if name + '/w' in params:
      return np.asarray(params[name + '/w'], dtype=np.float64), np.asarray(params[name + '/b'], dtype=np.float64)
    return None
  out, i = [], 0
  while True:
    layer = get('linear' if i == 0 else f'linear_{i}')
    if layer is None:
      break
    out.append(layer)
    i += 1
  if len(out) < 2:
    raise ValueError("expected haiku parameters 'linear', 'linear_1', ... (create_node.py:124-131)")
  return out


class InvasivePlant(FiniteHorizonControlSystem):
  """myriad/systems/lenhart/invasive_plant.py:38-101: the one DISCRETE system (x_{t+1} = f(x_t, u_t), five foci, one control
  each).  Like in the reference it works with the FBSM only -- the direct optimizers raise (trajectory_optimizers/base.py:66-67);
  its map, previous-adjoint rule and characterisation live in csrc/fbsm.cuh."""

  def __init__(self, B=1., k=1., eps=.01, x_0=(.5, 1., 1.5, 2., 10.), T=10.):
    inf = np.inf
    super().__init__(x_0=np.array(x_0, dtype=np.float64), x_T=None, T=T,
                     bounds=np.array([[-inf, inf]] * 5 + [[0., 1.]] * 5), terminal_cost=False, discrete=True,
                     device_name="INVASIVEPLANT", params=[B, k, eps])


class _NotOnDevice:
  def __init__(self, name):
    self.name = name

  def __call__(self, *a, **k):
    raise NotImplementedError(f"system {self.name} has no generated device code yet: add it to tools/gen_systems.py "
                              "(DESIGN.md, 'what comes next')")


class SystemType(Enum):
  """Same member names as the reference enum; members without a device implementation raise on call."""
  CARTPOLE = CartPole
  VANDERPOL = VanDerPol
  SEIR = SEIR
  TUMOUR = Tumour
  MOUNTAINCAR = MountainCar
  PENDULUM = Pendulum
  SIMPLECASE = SimpleCase
  MOULDFUNGICIDE = MouldFungicide
  BACTERIA = Bacteria
  SIMPLECASEWITHBOUNDS = SimpleCaseWithBounds
  CANCERTREATMENT = CancerTreatment
  EPIDEMICSEIRN = EpidemicSEIRN
  HARVEST = Harvest
  HIVTREATMENT = HIVTreatment
  BEARPOPULATIONS = BearPopulations
  GLUCOSE = Glucose
  TIMBERHARVEST = TimberHarvest
  BIOREACTOR = Bioreactor
  PREDATORPREY = PredatorPrey
  INVASIVEPLANT = InvasivePlant
  ROCKETLANDING = RocketLanding

  def __call__(self, *args, **kwargs) -> FiniteHorizonControlSystem:
    return self.value(*args, **kwargs)
