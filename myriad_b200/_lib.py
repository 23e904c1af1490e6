"""ctypes binding of libmyriad_b200.so (C ABI in include/myriad_b200.h).

The library is built in-tree by ``myriad_b200.build`` (``__graft_entry__.build()``).  There is no
fallback: if the shared object is missing, importing the product raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MYR_LIB", os.path.join(HERE, "libmyriad_b200.so"))  # MYR_LIB: A/B builds of the same ABI

ABI_VERSION = 4
WS_HEADER = 16  # MYR_WS_HEADER
MAX_PARAMS = 16
MAX_NODE_LAYERS = 5
NODE_BASE = 100

SYSTEM_IDS = {
  "SIMPLECASE": 0, "CARTPOLE": 1, "VANDERPOL": 2, "CANCERTREATMENT": 3, "MOULDFUNGICIDE": 4, "BIOREACTOR": 5,
  "SIMPLECASEWITHBOUNDS": 6, "GLUCOSE": 7, "HARVEST": 8, "TIMBERHARVEST": 9, "SEIR": 10, "EPIDEMICSEIRN": 11, "HIVTREATMENT": 12, "BACTERIA": 13, "TUMOUR": 14, "PREDATORPREY": 15, "BEARPOPULATIONS": 16,
  "ROCKETLANDING": 17, "PENDULUM": 18, "MOUNTAINCAR": 19, "INVASIVEPLANT": 20,
}
OPT_SHOOTING, OPT_TRAPEZOIDAL, OPT_HERMITE_SIMPSON = 0, 1, 2
METHOD_IDS = {"EULER": 0, "HEUN": 1, "MIDPOINT": 2, "RK4": 3}

STATUS_NAMES = {0: "solved", 1: "acceptable", -1: "max_iter", -2: "line_search_failed", -3: "inertia_correction_failed",
                -13: "invalid_number"}


class MyrDesc(C.Structure):
  _fields_ = [("system_id", C.c_int32), ("optimizer", C.c_int32), ("integration_method", C.c_int32),
              ("intervals", C.c_int32), ("controls_per_interval", C.c_int32), ("n_params", C.c_int32),
              ("terminal_cost", C.c_int32), ("reserved", C.c_int32), ("T", C.c_double),
              ("params", C.c_double * MAX_PARAMS),
              ("node_num_hidden", C.c_int32), ("node_hidden", C.c_int32 * (MAX_NODE_LAYERS - 1)),
              ("theta", C.c_void_p), ("theta_doubles", C.c_int64)]


class MyrSizes(C.Structure):
  _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("nx_nodes", C.c_int32), ("nu_nodes", C.c_int32),
              ("nvars", C.c_int32), ("ncon", C.c_int32), ("nodes", C.c_int32), ("stages", C.c_int32),
              ("nw", C.c_int32), ("nc", C.c_int32), ("stage_nodes", C.c_int32), ("ipm_workspace_slots", C.c_int32),
              ("jac_block_doubles", C.c_int64), ("hess_block_doubles", C.c_int64),
              ("ipm_workspace_doubles", C.c_int64)]


class MyrIpmOpts(C.Structure):
  _fields_ = [("max_iter", C.c_int32), ("max_ls", C.c_int32), ("acceptable_iter", C.c_int32), ("max_soc", C.c_int32),
              ("tol", C.c_double), ("acceptable_tol", C.c_double), ("mu_init", C.c_double)]


class MyrFbsmOpts(C.Structure):
  _fields_ = [("max_iter", C.c_int32), ("max_secant", C.c_int32), ("term_state", C.c_int32), ("reserved", C.c_int32),
              ("delta", C.c_double), ("secant_tol", C.c_double), ("term_value", C.c_double),
              ("guess_a", C.c_double), ("guess_b", C.c_double)]


class MyriadError(RuntimeError):
  pass


_P = C.c_void_p
_lib = None

EXPORTS = ["myr_abi_version", "myr_last_error", "myr_problem_sizes", "myr_eval", "myr_kkt_solve", "myr_ipm_solve",
           "myr_rollout_cost", "myr_dynamics", "myr_jtvec", "myr_register_system", "myr_bench_dfma", "myr_host_eval", "myr_host_kkt_solve", "myr_host_ipm_solve",
           "myr_host_rollout_cost", "myr_host_dynamics", "myr_host_jtvec", "myr_fbsm_solve", "myr_host_fbsm_solve"]


def lib() -> C.CDLL:
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise MyriadError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a).  myriad_b200 has no CPU fallback.")
  L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)   # plugin libraries (register_system) resolve myr::fail etc. against it
  L.myr_abi_version.restype = C.c_int
  L.myr_last_error.restype = C.c_char_p
  L.myr_problem_sizes.argtypes = [C.POINTER(MyrDesc), C.POINTER(MyrSizes)]
  ev = [C.POINTER(MyrDesc), C.c_int, _P, _P, _P, _P, _P, _P, _P]
  L.myr_eval.argtypes = ev + [_P]
  L.myr_host_eval.argtypes = ev
  kk = [C.POINTER(MyrDesc), C.c_int, _P, _P, _P, _P, _P, C.c_double, C.c_double, _P, _P, _P, _P, C.c_size_t]
  L.myr_kkt_solve.argtypes = kk + [_P]
  L.myr_host_kkt_solve.argtypes = kk
  ip = [C.POINTER(MyrDesc), C.POINTER(MyrIpmOpts), C.c_int] + [_P] * 12 + [_P, C.c_size_t]
  L.myr_ipm_solve.argtypes = ip + [_P]
  L.myr_host_ipm_solve.argtypes = ip
  ro = [C.POINTER(MyrDesc), C.c_int, C.c_int, _P, _P, _P, _P]
  L.myr_rollout_cost.argtypes = ro + [_P]
  L.myr_host_rollout_cost.argtypes = ro
  dy = [C.POINTER(MyrDesc), C.c_int, _P, _P, _P, _P, _P]
  L.myr_dynamics.argtypes = dy + [_P]
  L.myr_host_dynamics.argtypes = dy
  jt = [C.POINTER(MyrDesc), C.c_int, _P, _P, _P]
  L.myr_jtvec.argtypes = jt + [_P]
  L.myr_host_jtvec.argtypes = jt
  fb = [C.POINTER(MyrDesc), C.POINTER(MyrFbsmOpts), C.c_int] + [_P] * 9
  L.myr_fbsm_solve.argtypes = fb + [_P]
  L.myr_host_fbsm_solve.argtypes = fb
  L.myr_register_system.argtypes = [_P]
  L.myr_bench_dfma.argtypes = [C.c_int, C.c_int, _P, _P]
  for name in EXPORTS:
    if name not in ("myr_abi_version", "myr_last_error"):
      getattr(L, name).restype = C.c_int
  _lib = L
  return L


def check(rc: int) -> None:
  if rc != 0:
    msg = lib().myr_last_error().decode()
    if rc == -1:
      raise KeyError(msg)  # the reference raises KeyError / ValueError on unknown enums
    if rc == -2:
      raise NotImplementedError(msg)
    raise MyriadError(f"myriad_b200 error {rc}: {msg}")


USER_BASE = 1000  # MYR_SYS_USER_BASE


def make_desc(system: str, optimizer: int, method: str, intervals: int, cpi: int = 1, T: float = 0.0, params=None,
              terminal_cost: bool = False, hidden=None, theta_ptr: int = 0, theta_doubles: int = 0, keepalive=None) -> MyrDesc:
  """hidden / theta_ptr / theta_doubles: NODE systems ("NODE_<true system>") only; theta_ptr is a device pointer for the
  device entry points (host pointer for myr_host_*); ``keepalive`` (the tensor / array owning it) is pinned to the struct."""
  d = MyrDesc()
  node = system.startswith("NODE_")
  base = system[5:] if node else system
  if base not in SYSTEM_IDS:
    raise KeyError(f"system {system} has no device implementation")
  d.system_id = SYSTEM_IDS[base] + (NODE_BASE if node else 0)
  if node:
    if not hidden or len(hidden) > MAX_NODE_LAYERS - 1:
      raise KeyError(f"NODE systems need 1..{MAX_NODE_LAYERS - 1} hidden layers")
    d.node_num_hidden = len(hidden)
    for i, h in enumerate(hidden):
      d.node_hidden[i] = int(h)
    d.theta = int(theta_ptr)
    d.theta_doubles = int(theta_doubles)
    d._keepalive = keepalive
  d.optimizer = int(optimizer)
  d.integration_method = METHOD_IDS[method]
  d.intervals = int(intervals)
  d.controls_per_interval = int(cpi)
  d.T = float(T)
  d.terminal_cost = int(bool(terminal_cost))
  if params:
    d.n_params = len(params)
    for i, v in enumerate(params):
      d.params[i] = float(v)
  return d


def workspace_doubles(sizes: MyrSizes, B: int) -> int:
  """include/myriad_b200.h, MYR_WS_HEADER: a header plus one slot per resident CTA (never more than B)"""
  return WS_HEADER + max(1, min(int(B), int(sizes.ipm_workspace_slots))) * int(sizes.ipm_workspace_doubles)


def problem_sizes(desc: MyrDesc) -> MyrSizes:
  s = MyrSizes()
  check(lib().myr_problem_sizes(C.byref(desc), C.byref(s)))
  return s
