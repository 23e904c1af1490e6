"""Symbolic system -> device code generator (shared by tools/gen_systems.py, which writes csrc/systems_gen.cuh for the
built-in SystemType members, and by myriad_b200.plugin.register_system, which builds one extra translation unit for a
user-defined system).

A system is given as sympy expressions of its dynamics f(x, u, p), running cost g(x, u, t, p) and (optionally) a
terminal cost linear in x_T.  sympy differentiates them; common sub-expression elimination produces straight-line
fp64 code for f, [A|B] = df/d(x,u), sum_i mu_i hess f_i, g, grad g and hess g -- what the kernels consume.
"""
from __future__ import annotations

import sympy as sp
from sympy.printing.c import C99CodePrinter


class Printer(C99CodePrinter):
  def _print_Pow(self, expr):
    b, e = expr.base, expr.exp
    if e.is_Integer and 2 <= int(e) <= 4:
      s = self.parenthesize(b, 100)  # force parens unless atom
      return "(" + "*".join([s] * int(e)) + ")"
    if e.is_Integer and -4 <= int(e) <= -1:
      s = self.parenthesize(b, 100)
      return "(1.0/(" + "*".join([s] * (-int(e))) + "))"
    return super()._print_Pow(expr)

  def _print_Rational(self, expr):
    return f"({int(expr.p)}.0/{int(expr.q)}.0)"

  def _print_Integer(self, expr):
    return f"{int(expr)}.0"

  def _print_angnorm(self, expr):
    a = self._print(expr.args[0])
    # Python-style remainder (result has the sign of the divisor), like jnp.remainder
    return f"((({a}) + M_PI) - floor((({a}) + M_PI) / (2.0 * M_PI)) * (2.0 * M_PI) - M_PI)"


PR = Printer()


def cc(e):
  return PR.doprint(e)


def X(n):
  return sp.symbols(f"x0:{n}", real=True)


def U(m):
  return sp.symbols(f"u0:{m}", real=True)


t = sp.Symbol("t", real=True)


def clip(v, lo, hi):
  """jnp.clip(v, lo, hi) with JAX's sub-gradient choice: derivative 1 strictly inside, 0 outside"""
  return sp.Piecewise((lo, v < lo), (hi, v > hi), (v, True))


class angnorm(sp.Function):
  """angle_normalize(x) = ((x + pi) % (2 pi)) - pi (pendulum.py:17-18); jnp.remainder has derivative 1 w.r.t. x"""
  nargs = 1

  def fdiff(self, argindex=1):
    return sp.Integer(1)


def packed_index(i, j, nw):
  if i > j:
    i, j = j, i
  return i * nw - (i * (i - 1)) // 2 + (j - i)


def emit_block(lines, assigns, indent="    "):
  """CSE over all right-hand sides and print; assigns = [(lhs_string, expr, op)] op in {'=', '+='}"""
  exprs = [e for _, e, _ in assigns]
  repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("v"), optimizations="basic")
  # merge sin/cos pairs of the same argument into one sincos
  done = set()
  sc_pairs = {}
  for sym, ex in repl:
    if ex.func in (sp.sin, sp.cos):
      sc_pairs.setdefault(ex.args[0], {})[ex.func] = sym
  for sym, ex in repl:
    if ex.func in (sp.sin, sp.cos) and len(sc_pairs.get(ex.args[0], {})) == 2:
      arg = ex.args[0]
      if arg in done:
        continue
      done.add(arg)
      s_sym, c_sym = sc_pairs[arg][sp.sin], sc_pairs[arg][sp.cos]
      lines.append(f"{indent}double {s_sym}, {c_sym}; sincos({cc(arg)}, &{s_sym}, &{c_sym});")
    else:
      lines.append(f"{indent}const double {sym} = {cc(ex)};")
  for (lhs, _, op), ex in zip(assigns, red):
    if op == "+=" and ex == 0:
      continue
    lines.append(f"{indent}{lhs} {op} {cc(ex)};")


def gen_system(name, d):
  n, m = d["n"], d["m"]
  nw = n + m
  x, u = X(n), U(m)
  pn = [k for k, _ in d["params"]]
  psym = {k: sp.Symbol(f"p_{k}", real=True) for k in pn}
  f = [sp.sympify(e) for e in d["f"](x, u, psym)]
  g = sp.sympify(d["g"](x, u, t, psym))
  w = list(x) + list(u)
  mu = sp.symbols(f"mu0:{n}", real=True)
  wq = sp.Symbol("wq", real=True)
  J = [[sp.diff(fi, wj) for wj in w] for fi in f]
  L = sum(mu[i] * f[i] for i in range(n))
  Hf = [[sp.diff(L, w[i], w[j]) for j in range(nw)] for i in range(nw)]
  gg = [sp.diff(g, wj) for wj in w]
  Hg = [[sp.diff(g, w[i], w[j]) for j in range(nw)] for i in range(nw)]
  time_dep = g.has(t)

  cls = name.title().replace("_", "")
  out = []
  out.append(f"// {name}: {d['ref']}")
  out.append(f"struct Sys{cls} {{")
  out.append(f"  static constexpr int id = {d['id']}, n = {n}, m = {m}, nw = {nw}, np = {len(pn)};")
  out.append(f"  static constexpr bool time_dependent_cost = {'true' if time_dep else 'false'};")
  out.append(f"  static constexpr const char* name = \"{name}\";")
  out.append("  MYR_HD static void default_params(double* p) {")
  for i, (k, v) in enumerate(d["params"]):
    out.append(f"    p[{i}] = {v!r};  // {k}")
  out.append("  }")

  def prologue(lines, need_u=True, need_t=False):
    for i in range(n):
      lines.append(f"    const double x{i} = x[{i}];")
    for i in range(m):
      lines.append(f"    const double u{i} = u[{i}];")
    for i, k in enumerate(pn):
      lines.append(f"    const double p_{k} = p[{i}];")

  def fn(sig, body_assigns, ret=None, pre=None):
    lines = [f"  MYR_HD static {sig} {{"]
    prologue(lines)
    if pre:
      lines += pre
    emit_block(lines, body_assigns)
    if ret:
      lines.append(f"    return {ret};")
    lines.append("  }")
    # silence unused-variable warnings
    lines.insert(1, "    (void)x; (void)u; (void)p;")
    return lines

  # f
  out += fn("void f(const double* x, const double* u, const double* p, double* f)",
            [(f"f[{i}]", f[i], "=") for i in range(n)])
  # f + jac (row-major n x nw)
  out += fn("void fjac(const double* x, const double* u, const double* p, double* f, double* J)",
            [(f"f[{i}]", f[i], "=") for i in range(n)] +
            [(f"J[{i * nw + j}]", J[i][j], "=") for i in range(n) for j in range(nw)])
  # f + jac + H += sum mu_i hess f_i (packed upper, row-major)
  mu_pre = [f"    const double mu{i} = mu[{i}];" for i in range(n)]
  out += fn("void fjac_hess(const double* x, const double* u, const double* p, const double* mu, double* f, double* J, double* H)",
            [(f"f[{i}]", f[i], "=") for i in range(n)] +
            [(f"J[{i * nw + j}]", J[i][j], "=") for i in range(n) for j in range(nw)] +
            [(f"H[{packed_index(i, j, nw)}]", Hf[i][j], "+=") for i in range(nw) for j in range(i, nw)],
            pre=mu_pre)
  # cost
  out += fn("double cost(const double* x, const double* u, double t, const double* p)",
            [("const double g_", g, "=")], ret="g_", pre=["    (void)t;"])
  out += fn("double cost_grad(const double* x, const double* u, double t, const double* p, double* g)",
            [("const double g_", g, "=")] + [(f"g[{j}]", gg[j], "=") for j in range(nw)], ret="g_", pre=["    (void)t;"])
  out += fn("double cost_grad_hess(const double* x, const double* u, double t, const double* p, double wq, double* g, double* H)",
            [("const double g_", g, "=")] + [(f"g[{j}]", gg[j], "=") for j in range(nw)] +
            [(f"H[{packed_index(i, j, nw)}]", wq * Hg[i][j], "+=") for i in range(nw) for j in range(i, nw)],
            ret="g_", pre=["    (void)t; (void)wq;"])
  # terminal cost (systems/base.py:101-111): every reference system with terminal_cost=True has one that is LINEAR in
  # x_T and independent of u_T (bacteria.py:84-86, tumour.py:106-108, predator_prey.py:120-122), so the device code
  # carries only its coefficient vector: term(x_T) = sum_i tc[i] x_T[i]
  tc = [sp.Integer(0)] * n
  if d["term"] is not None:
    term = sp.sympify(d["term"](x, u, psym))
    tc = [sp.diff(term, xi) for xi in x]
    resid = sp.simplify(term - sum(c * xi for c, xi in zip(tc, x)))
    if resid != 0 or any(c.has(*x) or c.has(*u) for c in tc):
      raise NotImplementedError(f"{name}: only terminal costs linear in x_T are generated")
  out.append(f"  static constexpr bool has_terminal = {'true' if d['term'] is not None else 'false'};")
  lines_tc = ["  MYR_HD static void terminal_coef(const double* p, double* tc) {", "    (void)p;"]
  for i, k in enumerate(pn):
    lines_tc.append(f"    const double p_{k} = p[{i}];")
  for i in range(n):
    lines_tc.append(f"    tc[{i}] = {cc(tc[i])};")
  lines_tc.append("  }")
  out += lines_tc
  out.append("};")
  out.append("")
  return out


