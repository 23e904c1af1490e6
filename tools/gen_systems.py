#!/usr/bin/env python
"""Generate myriad_b200/csrc/systems_gen.cuh: per-system device functions (dynamics, running
cost, their Jacobians and multiplier-contracted Hessians) from symbolic definitions.

Each system below restates the formulas of the reference system it is named after (file:line
cited per entry; table in SURVEY.md section 2.4).  sympy differentiates them and common
sub-expression elimination produces straight-line fp64 code, so every kernel that needs
f, [A|B] = df/d(x,u) or sum_i mu_i * hess f_i gets them without autodiff at run time.

Run:  python tools/gen_systems.py   (writes the header; the header is committed)

To add a system (the reference's "System plugin" extension point, myriad/systems/__init__.py:29-50):
add an entry to SYSTEMS here, re-run, rebuild (python -c "import __graft_entry__ as g; g.build()"),
and add the matching descriptor in myriad_b200/systems/__init__.py.
"""
from __future__ import annotations

import os
import sys

import sympy as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "myriad_b200", "csrc", "systems_gen.cuh")


from myriad_b200.codegen import PR, U, X, angnorm, cc, clip, emit_block, gen_system, packed_index, t  # noqa: E402,F401


def system_defs():
  """name -> dict(id, n, m, params [(name, default)], f(x,u,p), g(x,u,t,p), term(x,u,p) or None, ref)"""
  S = {}

  def add(name, sid, n, m, params, f, g, term=None, ref=""):
    S[name] = dict(id=sid, n=n, m=m, params=params, f=f, g=g, term=term, ref=ref)

  # myriad/systems/lenhart/simple_case.py:46-53
  add("SIMPLECASE", 0, 1, 1, [("A", 1.0), ("B", 1.0), ("C", 4.0)],
      lambda x, u, p: [-sp.Rational(1, 2) * x[0] ** 2 + p["C"] * u[0]],
      lambda x, u, t, p: -p["A"] * x[0] + p["B"] * u[0] ** 2,
      ref="myriad/systems/lenhart/simple_case.py:46-53")

  # myriad/systems/classical_control/cartpole.py:76-87,106-108 (code, not docstring: SURVEY 9-1)
  def cartpole_f(x, u, p):
    g_, m1, m2, l = p["g"], p["m1"], p["m2"], p["length"]
    th, dx, dth = x[1], x[2], x[3]
    s, c = sp.sin(th), sp.cos(th)
    ddx = (l * m2 * s * dth ** 2 + u[0] + m2 * g_ * c * s) / (m1 + m2 * (1 - c ** 2))
    ddth = -(l * m2 * c * dth ** 2 + u[0] * c + (m1 + m2) * g_ * s) / (l * m1 + l * m2 * (1 - c ** 2))
    return [dx, dth, ddx, ddth]

  add("CARTPOLE", 1, 4, 1, [("g", 9.81), ("m1", 1.0), ("m2", 0.3), ("length", 0.5)],
      cartpole_f, lambda x, u, t, p: u[0] ** 2,
      ref="myriad/systems/classical_control/cartpole.py:76-87,106-108")

  # myriad/systems/miscellaneous/van_der_pol.py:46-60
  add("VANDERPOL", 2, 2, 1, [("a", 1.0)],
      lambda x, u, p: [p["a"] * (1 - x[1] ** 2) * x[0] - x[1] + u[0], x[0]],
      lambda x, u, t, p: x[0] ** 2 + x[1] ** 2 + u[0] ** 2,
      ref="myriad/systems/miscellaneous/van_der_pol.py:46-60")

  # myriad/systems/lenhart/cancer_treatment.py:62-76
  add("CANCERTREATMENT", 3, 1, 1, [("r", 0.3), ("a", 3.0), ("delta", 0.45)],
      lambda x, u, p: [p["r"] * x[0] * sp.log(1 / x[0]) - u[0] * p["delta"] * x[0]],
      lambda x, u, t, p: p["a"] * x[0] ** 2 + u[0] ** 2,
      ref="myriad/systems/lenhart/cancer_treatment.py:62-76")

  # ---- further reference systems, smooth ones only
  # myriad/systems/lenhart/mould_fungicide.py:50-66
  add("MOULDFUNGICIDE", 4, 1, 1, [("r", 0.3), ("M", 10.0), ("A", 10.0)],
      lambda x, u, p: [p["r"] * (p["M"] - x[0]) - u[0] * x[0]],
      lambda x, u, t, p: p["A"] * x[0] ** 2 + u[0] ** 2,
      ref="myriad/systems/lenhart/mould_fungicide.py:50-66")
  # myriad/systems/lenhart/bioreactor.py:61-83
  add("BIOREACTOR", 5, 1, 1, [("K", 2.0), ("G", 1.0), ("D", 1.0)],
      lambda x, u, p: [p["G"] * u[0] * x[0] - p["D"] * x[0] ** 2],
      lambda x, u, t, p: -p["K"] * x[0] + u[0],
      ref="myriad/systems/lenhart/bioreactor.py:61-83")
  # myriad/systems/lenhart/simple_case_with_bounds.py:49-56
  add("SIMPLECASEWITHBOUNDS", 6, 1, 1, [("A", 1.0), ("C", 4.0)],
      lambda x, u, p: [-sp.Rational(1, 2) * x[0] ** 2 + p["C"] * u[0]],
      lambda x, u, t, p: -p["A"] * x[0] + u[0] ** 2,
      ref="myriad/systems/lenhart/simple_case_with_bounds.py:49-56")
  # myriad/systems/lenhart/glucose.py:72-104
  add("GLUCOSE", 7, 2, 1, [("a", 1.0), ("b", 1.0), ("c", 1.0), ("A", 2.0), ("l", 0.5)],
      lambda x, u, p: [-p["a"] * x[0] - p["b"] * x[1], -p["c"] * x[1] + u[0]],
      lambda x, u, t, p: 100000 * (p["A"] * (x[0] - p["l"]) ** 2 + u[0] ** 2),
      ref="myriad/systems/lenhart/glucose.py:72-104")
  # myriad/systems/lenhart/harvest.py:55-62 (time-dependent cost)
  add("HARVEST", 8, 1, 1, [("A", 5.0), ("k", 10.0), ("m", 0.2)],
      lambda x, u, p: [-(p["m"] + u[0]) * x[0]],
      lambda x, u, t, p: -p["A"] * (p["k"] * t / (t + 1)) * x[0] * u[0] + u[0] ** 2,
      ref="myriad/systems/lenhart/harvest.py:55-62")
  # myriad/systems/lenhart/timber_harvest.py:62-85 (time-dependent cost)
  add("TIMBERHARVEST", 9, 1, 1, [("r", 0.0), ("k", 1.0)],
      lambda x, u, p: [p["k"] * x[0] * u[0]],
      lambda x, u, t, p: -sp.exp(-p["r"] * t) * x[0] * (1 - u[0]),
      ref="myriad/systems/lenhart/timber_harvest.py:62-85")

  # myriad/systems/miscellaneous/seir.py:47-95 (state = S, E, I, N)
  def seir_f(x, u, p):
    S_, E_, I_, N_ = x
    return [p["b"] * N_ - p["d"] * S_ - p["c"] * S_ * I_ - u[0] * S_,
            p["c"] * S_ * I_ - (p["e"] + p["d"]) * E_,
            p["e"] * E_ - (p["g"] + p["a"] + p["d"]) * I_,
            (p["b"] - p["d"]) * N_ - p["a"] * I_]
  seir_params = [("A", 0.1), ("b", 0.525), ("d", 0.5), ("c", 0.0001), ("e", 0.5), ("g", 0.1), ("a", 0.2)]
  add("SEIR", 10, 4, 1, seir_params, seir_f, lambda x, u, t, p: p["A"] * x[2] + u[0] ** 2,
      ref="myriad/systems/miscellaneous/seir.py:83-95")

  # myriad/systems/lenhart/epidemic_seirn.py:81-95 (same vector field, different bounds / constructor)
  add("EPIDEMICSEIRN", 11, 4, 1, seir_params, seir_f, lambda x, u, t, p: p["A"] * x[2] + u[0] ** 2,
      ref="myriad/systems/lenhart/epidemic_seirn.py:81-95")

  # myriad/systems/lenhart/hiv_treatment.py:74-111
  add("HIVTREATMENT", 12, 3, 1,
      [("s", 10.0), ("m_1", 0.02), ("m_2", 0.5), ("m_3", 4.4), ("r", 0.03), ("T_max", 1500.0), ("k", 0.000024), ("N", 300.0), ("A", 0.05)],
      lambda x, u, p: [p["s"] / (1 + x[2]) - p["m_1"] * x[0] + p["r"] * x[0] * (1 - (x[0] + x[1]) / p["T_max"]) - u[0] * p["k"] * x[0] * x[2],
                       u[0] * p["k"] * x[0] * x[2] - p["m_2"] * x[1],
                       p["N"] * p["m_2"] * x[1] - p["m_3"] * x[2]],
      lambda x, u, t, p: -p["A"] * x[0] + (1 - u[0]) ** 2,
      ref="myriad/systems/lenhart/hiv_treatment.py:74-111")
  # myriad/systems/lenhart/bacteria.py:34-86 (terminal cost -C x(T))
  add("BACTERIA", 13, 1, 1, [("r", 1.0), ("A", 1.0), ("B", 12.0), ("C", 1.0)],
      lambda x, u, p: [p["r"] * x[0] + p["A"] * u[0] * x[0] - p["B"] * u[0] ** 2 * sp.exp(-x[0])],
      lambda x, u, t, p: u[0] ** 2,
      term=lambda x, u, p: -p["C"] * x[0],
      ref="myriad/systems/lenhart/bacteria.py:59-86")

  # myriad/systems/miscellaneous/tumour.py:44-108 (no running cost, terminal cost p(T))
  add("TUMOUR", 14, 3, 1, [("xi", 0.084), ("b", 5.85), ("d", 0.00873), ("G", 0.15), ("mu", 0.02)],
      lambda x, u, p: [-p["xi"] * x[0] * sp.log(x[0] / x[1]),
                       x[1] * (p["b"] - (p["mu"] + p["d"] * x[0] ** sp.Rational(2, 3) + p["G"] * u[0])),
                       u[0]],
      lambda x, u, t, p: sp.Integer(0),
      term=lambda x, u, p: x[0],
      ref="myriad/systems/miscellaneous/tumour.py:77-108")
  # myriad/systems/lenhart/predator_prey.py:47-122 (terminal cost x_0(T); x_T = [None, None, B])
  add("PREDATORPREY", 15, 3, 1, [("d_1", 0.1), ("d_2", 0.1), ("A", 1.0)],
      lambda x, u, p: [(1 - x[1]) * x[0] - p["d_1"] * x[0] * u[0], (x[0] - 1) * x[1] - p["d_2"] * x[1] * u[0], u[0]],
      lambda x, u, t, p: p["A"] * sp.Rational(1, 2) * u[0] ** 2,
      term=lambda x, u, p: x[0],
      ref="myriad/systems/lenhart/predator_prey.py:82-122")
  # myriad/systems/lenhart/bear_populations.py:36-110 (two controls)
  def bear_f(x, u, p):
    k = p["r"] / p["K"]
    k2 = p["r"] / p["K"] ** 2
    return [p["r"] * x[0] - k * x[0] ** 2 + k * p["m_f"] * (1 - x[0] / p["K"]) * x[1] ** 2 - u[0] * x[0],
            p["r"] * x[1] - k * x[1] ** 2 + k * p["m_p"] * (1 - x[1] / p["K"]) * x[0] ** 2 - u[1] * x[1],
            k * (1 - p["m_p"]) * x[0] ** 2 + k * (1 - p["m_f"]) * x[1] ** 2 + k2 * p["m_f"] * x[0] * x[1] ** 2
            + k2 * p["m_p"] * (x[0] ** 2) * x[1]]
  add("BEARPOPULATIONS", 16, 3, 2, [("r", 0.1), ("K", 0.75), ("m_p", 0.5), ("m_f", 0.5), ("c_p", 10000.0), ("c_f", 10.0)],
      bear_f, lambda x, u, t, p: x[2] + p["c_p"] * u[0] ** 2 + p["c_f"] * u[1] ** 2,
      ref="myriad/systems/lenhart/bear_populations.py:71-110")
  # myriad/systems/miscellaneous/rocket_landing.py:100-124 (n = 6, two controls).  max_thrust and the rod inertia
  # I = m l^2 / 12 are derived in the constructor (:56-62): max_thrust is a parameter here so that the descriptor carries it
  def rocket_f(x, u, p):
    th = x[4]
    thrust, ang = u[0], u[1]
    Fx = p["max_thrust"] * thrust * sp.sin(ang + th)
    Fy = p["max_thrust"] * thrust * sp.cos(ang + th)
    Tq = -p["length"] / 2 * p["max_thrust"] * thrust * sp.sin(ang)
    I = sp.Rational(1, 12) * p["m"] * p["length"] ** 2
    return [x[1], Fx / p["m"], x[3], Fy / p["m"] - p["g"], x[5], Tq / I]
  add("ROCKETLANDING", 17, 6, 2, [("g", 9.8), ("m", 100000.0), ("length", 50.0), ("max_thrust", 2210000.0)],
      rocket_f, lambda x, u, t, p: u[0] ** 2 + u[1] ** 2 + 2 * x[5] ** 2,
      ref="myriad/systems/miscellaneous/rocket_landing.py:100-124")

  # myriad/systems/classical_control/pendulum.py:96-119: clipped torque and speed, normalised angle (non-smooth)
  def pendulum_f(x, u, p):
    uu = clip(u[0], -p["max_torque"], p["max_torque"])
    th = angnorm(x[0])
    dth = clip(x[1], -p["max_speed"], p["max_speed"])
    return [dth, (-3 * p["g"] / (2 * p["length"]) * sp.sin(th) + 3 * uu / (p["m"] * p["length"] ** 2)) * sp.Rational(1, 20)]
  add("PENDULUM", 18, 2, 1, [("g", 10.0), ("m", 1.0), ("length", 1.0), ("max_speed", 8.0), ("max_torque", 2.0), ("ctrl_penalty", 0.001)],
      pendulum_f, lambda x, u, t, p: angnorm(x[0]) ** 2 + sp.Rational(1, 10) * x[1] ** 2 + p["ctrl_penalty"] * u[0] ** 2,
      ref="myriad/systems/classical_control/pendulum.py:96-119")

  # myriad/systems/classical_control/mountain_car.py:86-107: hill_function(x) = x^2 / 2 (:11-13), so grad(hill)(x) = x
  add("MOUNTAINCAR", 19, 2, 1, [("power", 0.0015), ("gravity", 0.0025)],
      lambda x, u, p: [x[1], clip(u[0], -1, 1) * p["power"] - p["gravity"] * x[0]],
      lambda x, u, t, p: 10 * u[0] ** 2,
      ref="myriad/systems/classical_control/mountain_car.py:86-107")
  return S


def main():
  S = system_defs()
  lines = [
    "// GENERATED by tools/gen_systems.py -- do not edit by hand.",
    "// Per-system device functions: dynamics f(x,u), Jacobian J = df/d(x,u) (row-major n x (n+m)),",
    "// multiplier-contracted Hessian H += sum_i mu_i * d2 f_i / d(x,u)^2 (packed upper triangle,",
    "// row-major: idx(i,j) = i*nw - i*(i-1)/2 + (j-i)), running cost g(x,u,t) with gradient/Hessian.",
    "#pragma once",
    "#include <math.h>",
    "#ifndef MYR_HD",
    "#ifdef __CUDACC__",
    "#define MYR_HD __host__ __device__ __forceinline__",
    "#else",
    "#define MYR_HD inline",
    "#endif",
    "#endif",
    "",
    "namespace myr {",
    "",
  ]
  for name, d in S.items():
    lines += gen_system(name, d)
  # dispatch macro over all generated systems
  lines.append("#define MYR_FOR_EACH_SYSTEM(X) \\")
  names = list(S.keys())
  for i, name in enumerate(names):
    cls = name.title().replace("_", "")
    lines.append(f"  X(Sys{cls})" + (" \\" if i + 1 < len(names) else ""))
  lines.append("")
  lines.append("}  // namespace myr")
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  with open(OUT, "w") as fh:
    fh.write("\n".join(lines) + "\n")
  print("wrote", OUT, len(lines), "lines")


if __name__ == "__main__":
  sys.exit(main())
