"""refshim -- run the UNMODIFIED reference (/root/reference, nikihowe/myriad) without jax.

TEST INFRASTRUCTURE ONLY.  Nothing under ``myriad_b200/`` imports this.

The reference is pure Python on top of ``jax`` / ``cyipopt`` / ``gin`` / ``haiku`` ...,
none of which are installed in this image (no network).  This package installs tiny
stand-ins for those modules into ``sys.modules`` so that the reference's *own* Python
logic -- systems (``myriad/systems/*``), integrators (``myriad/utils.py:22-134``), the
three transcriptions (``myriad/trajectory_optimizers/*``) and ``myriad.nlp_solvers.solve``
with the SciPy solvers it supports (``myriad/nlp_solvers/__init__.py:50-55``) -- can be
executed here and its outputs frozen as golden fixtures (``oracle/make_golden.py``).

What the stand-ins are:

* ``jax.numpy``      -> NumPy (fp64; ``jax_enable_x64`` is what the reference sets, run.py:15)
* ``jax.jit``        -> identity
* ``jax.vmap``       -> Python loop over the mapped axis (honours ``in_axes``)
* ``jax.lax.scan``   -> Python loop
* ``jax.grad`` / ``jax.jacrev`` -> complex-step differentiation (exact to rounding for the
  analytic functions on this path; step 1e-30)
* ``jax.flatten_util.ravel_pytree`` -> concatenate + unravel closure (tuple/list of arrays)
* ``jax.random``     -> dummy keys (the hot path never draws from them: shooting.py:43-51)
* ``cyipopt`` / ``gin`` / ``tensorboardX`` / ``matplotlib`` / ``seaborn`` / ``haiku`` /
  ``optax`` / ``simple_parsing`` -> inert stubs (IPOPT itself is NOT available: calling
  ``minimize_ipopt`` raises).

So what gets pinned by the fixtures is the reference's Python logic evaluated in IEEE
fp64 by NumPy, not XLA's instruction selection; differences are at rounding level.
"""
from __future__ import annotations

import sys
import types

import numpy as np

REFERENCE_PATH = "/root/reference"


# ----------------------------------------------------------------------------- jnp
class ClampArray(np.ndarray):
  """ndarray whose integer indexing clamps past-the-end indices to the last element, like
  JAX's gather does (the reference relies on it: SURVEY.md section 9-17)."""

  def __getitem__(self, idx):
    if isinstance(idx, (int, np.integer)) and self.ndim >= 1 and idx >= self.shape[0]:
      idx = self.shape[0] - 1
    return super().__getitem__(idx)

  @property
  def at(self):
    """jnp's functional update ``a.at[idx].set(v)`` (forward_backward_sweep.py:89): returns an updated copy"""
    arr = self

    class _At:
      def __getitem__(self, idx):
        class _Setter:
          def set(self, v):
            out = np.array(arr)
            out[idx] = v
            return _wrap(out)
        return _Setter()
    return _At()

  def __iter__(self):  # iteration must still stop at the end
    for i in range(self.shape[0]):
      v = super().__getitem__(i)
      # complex scalars (complex-step differentiation) stay ClampArrays so that ``%`` below keeps working on them
      yield np.asarray(v).view(ClampArray) if np.iscomplexobj(v) and np.ndim(v) == 0 else v

  def __mod__(self, other):
    """jnp.remainder; for complex-step inputs the remainder acts on the real part (its derivative w.r.t. x is 1)"""
    if np.iscomplexobj(self):
      return _wrap(np.asarray(self) - np.floor(np.real(np.asarray(self)) / other) * other)
    return _wrap(np.mod(np.asarray(self), other))


def _wrap(v):
  if isinstance(v, np.ndarray) and not isinstance(v, ClampArray):
    return v.view(ClampArray)
  if isinstance(v, tuple):
    return tuple(_wrap(e) for e in v)
  return v


def _make_jnp() -> types.ModuleType:
  jnp = types.ModuleType("jax.numpy")

  def _getattr(name):
    if name == "NINF":
      return -np.inf
    attr = getattr(np, name)
    if callable(attr) and not isinstance(attr, type):
      def wrapped(*a, **k):
        return _wrap(attr(*a, **k))
      return wrapped
    return attr

  jnp.__getattr__ = _getattr  # type: ignore[attr-defined]
  jnp.ndarray = np.ndarray
  jnp.float64 = np.float64
  jnp.newaxis = np.newaxis
  jnp.pi = np.pi
  jnp.inf = np.inf
  jnp.NINF = -np.inf

  def array(x, dtype=None):
    # jnp.array([a, b, c]) with 0-d/1-element pieces, possibly complex (complex-step).
    try:
      return _wrap(np.array(x, dtype=dtype))
    except (ValueError, TypeError):
      return _wrap(np.array([np.asarray(e).reshape(()) for e in x], dtype=dtype))

  def clip(x, a_min=None, a_max=None):
    """jnp.clip with JAX's derivative (1 strictly inside, 0 outside), also for complex-step inputs"""
    x = np.asarray(x)
    if not np.iscomplexobj(x):
      return _wrap(np.clip(x, a_min, a_max))
    out = x.copy()
    if a_min is not None:
      out = np.where(x.real < a_min, a_min + 0j, out)
    if a_max is not None:
      out = np.where(x.real > a_max, a_max + 0j, out)
    return _wrap(out)

  jnp.array = array
  jnp.asarray = array
  jnp.clip = clip
  return jnp


# ----------------------------------------------------------------------------- jax core
def _vmap(fun, in_axes=0, out_axes=0):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    if len(axes) != len(args):
      # the reference has arity bugs here (trapezoidal.py:204-206, hermite_simpson.py:274-276);
      # real jax raises as well.
      raise ValueError(f"vmap in_axes {len(axes)} != number of args {len(args)}")
    n = None
    for a, ax in zip(args, axes):
      if ax is not None:
        n = np.shape(a)[ax]
        break
    outs = []
    for i in range(n):
      call = [a if ax is None else _wrap(np.take(a, i, axis=ax)) for a, ax in zip(args, axes)]
      outs.append(fun(*call))
    if isinstance(outs[0], tuple):
      return tuple(np.stack([np.asarray(o[j]) for o in outs]) for j in range(len(outs[0])))
    return np.stack([np.asarray(o) for o in outs])

  return mapped


def _jit(fun=None, **_kw):
  if fun is None:
    return lambda f: f
  return fun


def _scan(f, init, xs, length=None):
  carry = init
  ys = []
  for x in np.asarray(xs):
    carry, y = f(_wrap(carry) if isinstance(carry, np.ndarray) else carry, x)
    ys.append(np.asarray(y))
  return carry, np.stack(ys)


_CS_STEP = 1e-30


def _jacfwd_complex(fun):
  """Complex-step Jacobian of fun: R^n -> R^m (or scalar)."""

  def jac(z):
    z = np.asarray(z, dtype=np.float64)
    cols = []
    for j in range(z.shape[0]):
      zc = z.astype(np.complex128)
      zc[j] += 1j * _CS_STEP
      cols.append(np.imag(np.asarray(fun(_wrap(zc)))) / _CS_STEP)
    return np.stack(cols, axis=-1)

  return jac


def _grad(fun, argnums=0):
  """jax.grad of a scalar function w.r.t. positional argument ``argnums`` (complex-step)"""

  def g(*args):
    z = args[argnums]
    rest = lambda v: fun(*[v if k == argnums else a for k, a in enumerate(args)])
    if np.iscomplexobj(z) and np.ndim(z) == 0:
      # a grad nested inside an outer complex-step differentiation (mountain_car.py:91: jax.grad(hill_function)):
      # central difference with a real step on the complex point -- exact up to rounding for the quadratic hill
      # function (:11-13), and it carries the outer imaginary perturbation through
      h = 1e-3 * max(1.0, abs(complex(z)))
      return _wrap(np.asarray((rest(z + h) - rest(z - h)) / (2 * h)))
    if np.ndim(z) == 0:
      zc = complex(float(z), _CS_STEP)
      return _wrap(np.asarray(np.imag(rest(zc)) / _CS_STEP))
    return np.asarray(_jacfwd_complex(rest)(z)).reshape(-1)

  return g


def _ravel_pytree(tree):
  leaves = [np.asarray(l) for l in tree]
  shapes = [l.shape for l in leaves]
  sizes = [int(np.prod(s)) for s in shapes]
  flat = np.concatenate([l.ravel() for l in leaves])

  def unravel(v):
    out, o = [], 0
    for s, n in zip(shapes, sizes):
      out.append(np.reshape(v[o:o + n], s))
      o += n
    return type(tree)(out) if isinstance(tree, (tuple, list)) else out

  return flat, unravel


class _Stub(types.ModuleType):
  """Module whose every attribute is an inert callable/decorator."""

  def __getattr__(self, name):
    if name.startswith("__"):
      raise AttributeError(name)

    def _inert(*a, **k):
      if len(a) == 1 and callable(a[0]) and not k:
        return a[0]  # used as a bare decorator (@gin.configurable)
      return _Stub(name)

    return _inert

  def __call__(self, *a, **k):
    return _Stub("call")


def install(reference_path: str = REFERENCE_PATH) -> None:
  """Install the stand-ins and put the reference on sys.path.  Idempotent."""
  if "jax" in sys.modules and getattr(sys.modules["jax"], "_is_refshim", False):
    return
  jnp = _make_jnp()
  jax = types.ModuleType("jax")
  jax._is_refshim = True
  jax.numpy = jnp
  jax.vmap = _vmap
  jax.jit = _jit
  jax.grad = _grad
  jax.jacrev = _jacfwd_complex
  jax.jacfwd = _jacfwd_complex

  lax = types.ModuleType("jax.lax")
  lax.scan = _scan
  jax.lax = lax

  flatten_util = types.ModuleType("jax.flatten_util")
  flatten_util.ravel_pytree = _ravel_pytree
  jax.flatten_util = flatten_util

  random = types.ModuleType("jax.random")
  random.PRNGKey = lambda seed: np.array([0, seed], dtype=np.uint32)
  random.split = lambda key, num=2: tuple(np.array([i + 1, int(key[1])], dtype=np.uint32) for i in range(num))

  def _no_draw(*a, **k):
    raise RuntimeError("refshim: the hot path must not draw jax random numbers")

  random.normal = _no_draw
  random.uniform = _no_draw
  jax.random = random

  class _Cfg:
    def update(self, *a, **k):
      pass

  cfgmod = types.ModuleType("jax.config")
  cfgmod.config = _Cfg()
  jax.config = _Cfg()

  nn = types.ModuleType("jax.nn")
  nn.sigmoid = lambda x: 1.0 / (1.0 + np.exp(-x))
  jax.nn = nn

  sys.modules.update({
    "jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.flatten_util": flatten_util,
    "jax.random": random, "jax.config": cfgmod, "jax.nn": nn,
  })

  cyipopt = types.ModuleType("cyipopt")

  def minimize_ipopt(*a, **k):
    raise RuntimeError("refshim: IPOPT (cyipopt) is not available in this image")

  cyipopt.minimize_ipopt = minimize_ipopt
  sys.modules["cyipopt"] = cyipopt

  for name in ("gin", "tensorboardX", "matplotlib", "matplotlib.pyplot", "matplotlib.offsetbox",
               "seaborn", "haiku", "optax", "simple_parsing"):
    sys.modules[name] = _Stub(name)
  sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
  sys.modules["matplotlib"].offsetbox = sys.modules["matplotlib.offsetbox"]

  if reference_path not in sys.path:
    sys.path.insert(0, reference_path)
