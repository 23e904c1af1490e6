"""TEST INFRASTRUCTURE -- CPU restatement (NumPy) of the reference's Forward-Backward Sweep Method.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it; the product never does.

Follows, line for line in meaning:
  * FBSM.solve / sequencesolver / reinitiate   myriad/trajectory_optimizers/forward_backward_sweep.py:75-158
  * integrate_fbsm (RK4, both directions)      myriad/utils.py:138-197
  * stopping_criterion                         myriad/trajectory_optimizers/base.py:128-141
  * dynamics / adj_ODE / optim_characterization of the 14 indirect Lenhart systems  myriad/systems/lenhart/*.py
Pinned: tests/test_fbsm.py checks it against tests/golden/fbsm_*.npz, which oracle/make_fbsm_golden.py produced by
running the unmodified reference under oracle/refshim.
"""
from __future__ import annotations

import numpy as np

clamp = lambda v, lo, hi: np.minimum(hi, np.maximum(lo, v))


def _bang(temp, lo, hi):
  M = max(abs(lo), abs(hi))
  return clamp(np.sign(temp) * 2 * M + M, lo, hi)


# name -> dict(T, x_0, bounds, adj_T, x_T, guess, p, f(x,u,t,p), adj(a,x,u,t,p), opt(a,x,t,p,bounds))
# x, a: (n,) ; u: (m,) ; returns arrays of matching length.  Formulas: myriad/systems/lenhart/<file>.py as cited.
INF = np.inf
SYSTEMS = {
  "SIMPLECASE": dict(  # simple_case.py:24-63
    T=1., x_0=[1.], bounds=[[-INF, INF], [-INF, INF]], p=dict(A=1., B=1., C=4.),
    f=lambda x, u, t, p: -0.5 * x ** 2 + p["C"] * u,
    adj=lambda a, x, u, t, p: -p["A"] + x * a,
    opt=lambda a, x, t, p, b: clamp((p["C"] * a[0]) / (2 * p["B"]), b[0, 0], b[0, 1])),
  "SIMPLECASEWITHBOUNDS": dict(  # simple_case_with_bounds.py:24-66
    T=1., x_0=[1.], bounds=[[0., 3.], [-1., 2.]], p=dict(A=1., C=4.),
    f=lambda x, u, t, p: -0.5 * x ** 2 + p["C"] * u,
    adj=lambda a, x, u, t, p: -p["A"] + x * a,
    opt=lambda a, x, t, p, b: clamp((p["C"] * a[0]) / 2, b[-1, 0], b[-1, 1])),
  "CANCERTREATMENT": dict(  # cancer_treatment.py:40-96
    T=20., x_0=[.975], bounds=[[1e-3, 1.], [0., 2.]], p=dict(r=.3, a=3., delta=.45),
    f=lambda x, u, t, p: p["r"] * x * np.log(1 / x) - u * p["delta"] * x,
    adj=lambda a, x, u, t, p: a * (p["r"] + p["delta"] * u - p["r"] * np.log(1 / x)) - 2 * p["a"] * x,
    opt=lambda a, x, t, p, b: clamp(0.5 * a[0] * p["delta"] * x[0], b[-1, 0], b[-1, 1])),
  "MOULDFUNGICIDE": dict(  # mould_fungicide.py:29-70
    T=5., x_0=[1.], bounds=[[0., 5.], [0., 5.]], p=dict(r=.3, M=10., A=10.),
    f=lambda x, u, t, p: p["r"] * (p["M"] - x) - u * x,
    adj=lambda a, x, u, t, p: a * (p["r"] + u) - 2 * p["A"] * x,
    opt=lambda a, x, t, p, b: clamp(0.5 * a[0] * x[0], b[-1, 0], b[-1, 1])),
  "BIOREACTOR": dict(  # bioreactor.py:38-93
    T=2., x_0=[.5], bounds=[[0., 1.], [0., 1.]], p=dict(K=2., G=1., D=1.),
    f=lambda x, u, t, p: p["G"] * u * x - p["D"] * x ** 2,
    adj=lambda a, x, u, t, p: -p["K"] - p["G"] * u * a + 2 * p["D"] * x * a,
    opt=lambda a, x, t, p, b: _bang(-1 + p["G"] * a[0] * x[0], b[-1, 0], b[-1, 1])),
  "GLUCOSE": dict(  # glucose.py:46-105
    T=.2, x_0=[.75, 0.], bounds=[[0., 1.], [0., 1.], [0., .01]], p=dict(a=1., b=1., c=1., A=2., l=.5),
    f=lambda x, u, t, p: np.array([-p["a"] * x[0] - p["b"] * x[1], -p["c"] * x[1] + u[0]]),
    adj=lambda a, x, u, t, p: np.array([-2 * p["A"] * (x[0] - p["l"]) + a[0] * p["a"], a[0] * p["b"] + a[1] * p["c"]]),
    opt=lambda a, x, t, p, b: -a[1] / 2),
  "HARVEST": dict(  # harvest.py:32-72
    T=10., x_0=[.4], bounds=[[-INF, INF], [0., 1.]], p=dict(A=5., k=10., m=.2),
    f=lambda x, u, t, p: -(p["m"] + u) * x,
    adj=lambda a, x, u, t, p: a * (p["m"] + u) - p["A"] * (p["k"] * t / (t + 1)) * u,
    opt=lambda a, x, t, p, b: clamp(0.5 * x[0] * (p["A"] * (p["k"] * t / (t + 1)) - a[0]), b[-1, 0], b[-1, 1])),
  "TIMBERHARVEST": dict(  # timber_harvest.py:43-87
    T=5., x_0=[100.], bounds=[[0., 20_000.], [0., 1.]], p=dict(r=0., k=1.),
    f=lambda x, u, t, p: p["k"] * x * u,
    adj=lambda a, x, u, t, p: u * (np.exp(-p["r"] * t) - p["k"] * a) - np.exp(-p["r"] * t),
    opt=lambda a, x, t, p, b: _bang(x[0] * (p["k"] * a[0] - np.exp(-p["r"] * t)), b[-1, 0], b[-1, 1])),
  "EPIDEMICSEIRN": dict(  # epidemic_seirn.py:47-112
    T=20., x_0=[1000., 100., 50., 1165.], bounds=[[-INF, INF]] * 4 + [[0., .9]],
    p=dict(A=.1, b=.525, d=.5, c=.0001, e=.5, g=.1, a=.2),
    f=lambda x, u, t, p: np.array([
      p["b"] * x[3] - p["d"] * x[0] - p["c"] * x[0] * x[2] - u[0] * x[0],
      p["c"] * x[0] * x[2] - (p["e"] + p["d"]) * x[1],
      p["e"] * x[1] - (p["g"] + p["a"] + p["d"]) * x[2],
      (p["b"] - p["d"]) * x[3] - p["a"] * x[2]]),
    adj=lambda a, x, u, t, p: np.array([
      a[0] * (p["d"] + p["c"] * x[2] + u[0]) - a[1] * p["c"] * x[2],
      a[1] * (p["e"] + p["d"]) - a[2] * p["e"],
      -p["A"] + a[0] * p["c"] * x[0] - a[1] * p["c"] * x[0] + a[2] * (p["g"] + p["a"] + p["d"]) + a[3] * p["a"],
      -p["b"] * a[0] + a[3] * (p["d"] - p["d"])]),
    opt=lambda a, x, t, p, b: clamp(a[0] * x[0] / 2, b[-1, 0], b[-1, 1])),
  "HIVTREATMENT": dict(  # hiv_treatment.py:37-129
    T=20., x_0=[800., .04, 1.5], bounds=[[0., 1600.], [0., 100.], [0., 100.], [0., 1.]],
    p=dict(s=10., m_1=.02, m_2=.5, m_3=4.4, r=.03, T_max=1500., k=.000024, N=300., A=.05),
    f=lambda x, u, t, p: np.array([
      p["s"] / (1 + x[2]) - p["m_1"] * x[0] + p["r"] * x[0] * (1 - (x[0] + x[1]) / p["T_max"]) - u[0] * p["k"] * x[0] * x[2],
      u[0] * p["k"] * x[0] * x[2] - p["m_2"] * x[1],
      p["N"] * p["m_2"] * x[1] - p["m_3"] * x[2]]),
    adj=lambda a, x, u, t, p: np.array([
      -p["A"] + a[0] * (p["m_1"] - p["r"] * (1 - (x[0] + x[1]) / p["T_max"]) + p["r"] * x[0] / p["T_max"]
                        + u[0] * p["k"] * x[2]) - a[1] * u[0] * p["k"] * x[2],
      a[0] * p["r"] * x[0] / p["T_max"] + a[1] * p["m_2"] - a[2] * p["N"] * p["m_2"],
      a[0] * (p["s"] / (1 + x[2]) ** 2 + u[0] * p["k"] * x[0]) - a[1] * u[0] * p["k"] * x[0] + a[2] * p["m_3"]]),
    opt=lambda a, x, t, p, b: clamp(1 + 0.5 * p["k"] * x[0] * x[2] * (a[1] - a[0]), b[-1, 0], b[-1, 1])),
  "BACTERIA": dict(  # bacteria.py:34-96
    T=1., x_0=[1.], bounds=[[0., 10.], [0., 2.]], adj_T=[1.], p=dict(r=1., A=1., B=12., C=1.),
    f=lambda x, u, t, p: p["r"] * x + p["A"] * u * x - p["B"] * u ** 2 * np.exp(-x),
    adj=lambda a, x, u, t, p: -a * (p["r"] + p["A"] * u + p["B"] * u ** 2 * np.exp(-x)),
    opt=lambda a, x, t, p, b: clamp(a[0] * p["A"] * x[0] / (2 * (1 + p["B"] * a[0] * np.exp(-x[0]))), b[-1, 0], b[-1, 1])),
  "PREDATORPREY": dict(  # predator_prey.py:47-137
    T=10., x_0=[10., 1., 0.], bounds=[[0., 11.], [0., 11.], [0., 5.], [0., 1.]], adj_T=[1., 0., 0.], x_T=[None, None, 5.],
    guess=(-.52, .5), p=dict(d_1=.1, d_2=.1, A=1.),
    f=lambda x, u, t, p: np.array([(1 - x[1]) * x[0] - p["d_1"] * x[0] * u[0], (x[0] - 1) * x[1] - p["d_2"] * x[1] * u[0], u[0]]),
    adj=lambda a, x, u, t, p: np.array([a[0] * (x[1] - 1 + p["d_1"] * u[0]) - a[1] * x[1],
                                        a[0] * x[0] + a[1] * (1 - x[0] + p["d_2"] * u[0]), 0.]),
    opt=lambda a, x, t, p, b: clamp((a[0] * p["d_1"] * x[0] + a[1] * p["d_2"] * x[1] - a[2]) / p["A"], b[-1, 0], b[-1, 1])),
  "BEARPOPULATIONS": dict(  # bear_populations.py:40-139
    T=25., x_0=[.4, .2, 0.], bounds=[[0., 2.], [0., 2.], [0., 2.], [0., .2], [0., .2]],
    p=dict(r=.1, K=.75, m_p=.5, m_f=.5, c_p=10_000., c_f=10.),
    f=None, adj=None, opt=None),
}


def _bear_f(x, u, t, p):  # bear_populations.py:68-82
  k = p["r"] / p["K"]
  k2 = p["r"] / p["K"] ** 2
  return np.array([
    p["r"] * x[0] - k * x[0] ** 2 + k * p["m_f"] * (1 - x[0] / p["K"]) * x[1] ** 2 - u[0] * x[0],
    p["r"] * x[1] - k * x[1] ** 2 + k * p["m_p"] * (1 - x[1] / p["K"]) * x[0] ** 2 - u[1] * x[1],
    k * (1 - p["m_p"]) * x[0] ** 2 + k * (1 - p["m_f"]) * x[1] ** 2
    + k2 * p["m_f"] * x[0] * x[1] ** 2 + k2 * p["m_p"] * x[0] ** 2 * x[1]])


def _bear_adj(a, x, u, t, p):  # bear_populations.py:112-127
  k = p["r"] / p["K"]
  k2 = p["r"] / p["K"] ** 2
  return np.array([
    a[0] * (2 * k * x[0] + k2 * p["m_f"] * x[1] ** 2 + u[0] - p["r"]) - a[1] * (2 * k * p["m_p"] * (1 - x[1] / p["K"]) * x[0])
    + a[2] * (2 * k * (p["m_p"] - 1) * x[0] - k2 * p["m_f"] * x[1] ** 2 - 2 * k2 * p["m_p"] * x[0] * x[1]),
    a[1] * (2 * k * x[1] + k2 * p["m_p"] * x[0] ** 2 + u[1] - p["r"]) - a[0] * (2 * k * p["m_f"] * (1 - x[0] / p["K"]) * x[1])
    + a[2] * (2 * k * (p["m_f"] - 1) * x[1] - 2 * k2 * p["m_f"] * x[0] * x[1] - k2 * p["m_p"] * x[0] ** 2),
    -1.])


def _bear_opt(a, x, t, p, b):  # bear_populations.py:129-139
  return np.array([clamp(a[0] * x[0] / (2 * p["c_p"]), b[-2, 0], b[-2, 1]), clamp(a[1] * x[1] / (2 * p["c_f"]), b[-1, 0], b[-1, 1])])


SYSTEMS["BEARPOPULATIONS"].update(f=_bear_f, adj=_bear_adj, opt=_bear_opt)

# invasive_plant.py:38-90 -- DISCRETE: f is the next state, adj the previous adjoint, opt one row of the shifted rule
SYSTEMS["INVASIVEPLANT"] = dict(
  T=10., x_0=[.5, 1., 1.5, 2., 10.], bounds=[[-INF, INF]] * 5 + [[0., 1.]] * 5, adj_T=[1.] * 5, discrete=True,
  p=dict(B=1., k=1., eps=.01),
  f=lambda x, u, t, p: (x + x * p["k"] / (p["eps"] + x)) * (1 - u),
  adj=lambda a, x, u, t, p: a * (1 - u) * (1 + p["eps"] * p["k"] / (p["eps"] + x) ** 2),
  opt=lambda a, x, t, p, b: clamp(0.5 * a / p["B"] * (x + x * p["k"] / (p["eps"] + x)), b[-1, 0], b[-1, 1]))


def _rk4(fn, y, a0, a1, b0, b1, t, h):
  """rk4_step of integrate_fbsm (utils.py:166-176): fn(y, a, b, t)"""
  am, bm = (a0 + a1) / 2, (b0 + b1) / 2
  k1 = fn(y, a0, b0, t)
  k2 = fn(y + h * k1 / 2, am, bm, t + h / 2)
  k3 = fn(y + h * k2 / 2, am, bm, t + h / 2)
  k4 = fn(y + h * k3, a1, b1, t + h)
  return y + (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


def solve(name: str, N: int, x_0=None, max_iter: int = 10000):
  """-> dict(x, u, adj, sweeps) for one start state, exactly the reference's control flow."""
  S = SYSTEMS[name]
  p, T = S["p"], S["T"]
  b = np.asarray(S["bounds"], dtype=np.float64)
  x_0 = np.asarray(S["x_0"] if x_0 is None else x_0, dtype=np.float64)
  n, m = x_0.shape[0], b.shape[0] - x_0.shape[0]
  discrete = bool(S.get("discrete"))
  if discrete:  # forward_backward_sweep.py:33-35
    N = int(T)
  h = T / N
  ts = np.linspace(0, T, N + 1)
  f = lambda x, u, _v, t: np.atleast_1d(S["f"](x, u, t, p))
  g = lambda a, x, u, t: np.atleast_1d(S["adj"](a, x, u, t, p))
  sweeps = [0]

  def fixed_point(x, u, adj):
    while True:
      old_x, old_u, old_adj = x.copy(), u.copy(), adj.copy()
      x = x.copy()
      adj = adj.copy()
      if discrete:  # utils.py:182-186: direct evaluation, no integration
        for i in range(N):
          x[i + 1] = f(x[i], u[i], None, ts[i])
        for i in range(N, 0, -1):
          adj[i - 1] = g(adj[i], x[i], u[i - 1], ts[i - 1])
        est = np.stack([np.atleast_1d(S["opt"](adj[i + 1], x[i], ts[i], p, b)) for i in range(N)])  # shifted rows
      else:
        for i in range(N):  # forward (utils.py:190-192)
          x[i + 1] = _rk4(f, x[i], u[i], u[i + 1], 0.0, 0.0, ts[i], h)
        for i in range(N, 0, -1):  # backward (utils.py:194-196)
          adj[i - 1] = _rk4(g, adj[i], x[i], x[i - 1], u[i], u[i - 1], ts[i], -h)
        est = np.stack([np.atleast_1d(S["opt"](adj[i], x[i], ts[i], p, b)) for i in range(N + 1)])
      u = 0.5 * (est + old_u)
      sweeps[0] += 1
      sx = np.abs(x).sum(0) * 1e-3 - np.abs(x - old_x).sum(0)
      su = np.abs(u).sum(0) * 1e-3 - np.abs(u - old_u).sum(0)
      sa = np.abs(adj).sum(0) * 1e-3 - np.abs(adj - old_adj).sum(0)
      if not (np.min(np.hstack((su, sx, sa))) < 0) or sweeps[0] >= max_iter:
        return x, u, adj

  def guesses(a=None, ts_idx=None):
    x = np.vstack((x_0, np.zeros((N, n))))
    u = np.zeros((N if discrete else N + 1, m))
    adj = np.zeros((N + 1, n))
    if S.get("adj_T") is not None:
      adj[-1] = S["adj_T"]
    if a is not None:
      adj[-1, ts_idx] = a
    return x, u, adj

  x_T = S.get("x_T")
  if x_T is None:
    x, u, adj = fixed_point(*guesses())
  else:  # sequencesolver (forward_backward_sweep.py:118-158)
    idx = [i for i, v in enumerate(x_T) if v is not None][0]
    val = x_T[idx]
    a, c = S["guess"]
    x, u, adj = fixed_point(*guesses(a, idx)); Va = x[-1, idx] - val
    x, u, adj = fixed_point(*guesses(c, idx)); Vc = x[-1, idx] - val
    while abs(Va) > 1e-10:
      if abs(Va) > abs(Vc):
        a, c = c, a
        Va, Vc = Vc, Va
      d = Va * (c - a) / (Vc - Va)
      c, Vc = a, Va
      a = a - d
      x, u, adj = fixed_point(*guesses(a, idx)); Va = x[-1, idx] - val
  return dict(x=x, u=u, adj=adj, sweeps=sweeps[0])
