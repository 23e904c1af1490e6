"""Forward-Backward Sweep Method (myriad/trajectory_optimizers/forward_backward_sweep.py) parity.

tests/golden/fbsm_*.npz hold the UNMODIFIED reference's {'x', 'u', 'adj'} and its sweep counts for the 14 indirect (Lenhart)
systems, the discrete INVASIVEPLANT included (oracle/make_fbsm_golden.py, reference run under oracle/refshim).  CPU tests pin the NumPy restatement
(oracle/fbsm.py) and the host build of the sweep templates to them; the GPU tests run the product path
(get_optimizer(...).solve_batch -> myr_fbsm_solve) against the fixtures and, for random start states, against the oracle.

Tolerance (floating point): trajectories agree to 1e-9 relative to the trajectory's magnitude -- the kernel performs the
reference's operations in the reference's order, differences come from FMA contraction and the summation order of the
stopping rule -- and the number of sweeps must be identical.
"""
import os

import numpy as np
import pytest
import torch

from myriad_b200.config import Config, HParams, OptimizerType
from myriad_b200.systems import SystemType
from myriad_b200.trajectory_optimizers import get_optimizer
from oracle import fbsm as OF

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SYSTEMS = sorted(OF.SYSTEMS)
RTOL = 1e-9


def _gold(name):
  return np.load(os.path.join(GOLD, f"fbsm_{name.lower()}.npz"))


def _optimizer(name, N):
  hp = HParams(system=SystemType[name], optimizer=OptimizerType.FBSM, fbsm_intervals=N)
  return get_optimizer(hp, Config(verbose=False, plot=False), hp.system())


def _close(got, want, what):
  scale = max(1.0, float(np.abs(want).max()))
  err = float(np.abs(np.asarray(got) - want).max()) / scale
  assert err <= RTOL, f"{what}: {err:.3e}"


@pytest.mark.parametrize("name", SYSTEMS)
def test_oracle_restatement_matches_reference_fixture(name):
  g = _gold(name)
  N = int(g["N"])
  rows = range(len(g["x0"])) if N <= 200 else [0]
  for r in rows:
    s = OF.solve(name, N, g["x0"][r])
    assert s["sweeps"] == int(g["sweeps"][r])
    for k in ("x", "u", "adj"):
      _close(s[k], g[k][r], f"{name} row {r} {k}")


def test_oracle_system_data_matches_product_systems():
  for name in SYSTEMS:
    s = SystemType[name].value()
    S = OF.SYSTEMS[name]
    assert np.array_equal(np.asarray(s.x_0, dtype=np.float64), np.asarray(S["x_0"], dtype=np.float64)), name
    assert np.array_equal(np.asarray(s.bounds, dtype=np.float64), np.asarray(S["bounds"], dtype=np.float64)), name
    assert float(s.T) == float(S["T"]), name


@pytest.mark.parametrize("name", SYSTEMS)
def test_host_build_of_the_sweep_templates_matches_reference_fixture(name):
  g = _gold(name)
  opt = _optimizer(name, int(g["N"]))
  r = opt.host_solve_batch(g["x0"])
  assert (r["status"] == 0).all()
  assert r["iters"].tolist() == g["sweeps"].tolist()
  for k in ("x", "u", "adj"):
    _close(r[k], g[k], f"{name} {k}")


def test_fbsm_surface_mirrors_reference():
  opt = _optimizer("PREDATORPREY", 50)
  assert opt.require_adj and opt.terminal_cdtion and opt.term_cdtion_state == 2 and opt.term_value == 5.0
  assert opt.x_guess.shape == (51, 3) and opt.u_guess.shape == (51, 1) and opt.adj_guess.shape == (51, 3)
  assert opt.adj_guess[-1].tolist() == [1.0, 0.0, 0.0] and opt.t_interval.shape == (51, 1)
  assert opt.guess.shape == (51 * 7,)
  x = np.ones((5, 1))
  assert opt.stopping_criterion((x, x * 1.01), (x, x), (x, x)) is True
  assert opt.stopping_criterion((x, x * 1.0001), (x, x), (x, x)) is False
  hp = HParams(system=SystemType.CARTPOLE, optimizer=OptimizerType.FBSM)
  with pytest.raises(NotImplementedError):
    get_optimizer(hp, Config(verbose=False, plot=False), hp.system())


def test_host_empty_batch_and_unsupported_system():
  import ctypes as C
  from myriad_b200 import _lib as ML
  opt = _optimizer("SIMPLECASE", 10)
  r = opt.host_solve_batch(np.zeros((0, 1)))
  assert r["x"].shape == (0, 11, 1) and r["iters"].shape == (0,)
  d = ML.make_desc("CARTPOLE", ML.OPT_SHOOTING, "RK4", 10, 1)
  z = np.zeros(64)
  p = z.ctypes.data_as(C.c_void_p)
  rc = ML.lib().myr_host_fbsm_solve(C.byref(d), None, 1, p, None, p, p, p, p, p, p, p)
  assert rc == -2 and b"adj_ODE" in ML.lib().myr_last_error()


def _gloo_fbsm_worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  import torch.distributed as dist
  dist.init_process_group("gloo", rank=rank, world_size=world)
  opt = _optimizer("CANCERTREATMENT", 60)
  r = opt.solve_batch_sharded(_five_starts(), host=True)  # UNEQUAL shards: 3 + 2
  if rank == 0:
    q.put({k: v.numpy() for k, v in r.items()})
  dist.barrier()
  dist.destroy_process_group()


def _five_starts():
  return 0.975 * (1.0 + 0.05 * np.linspace(-1, 1, 5))[:, None]


def test_two_rank_sharded_fbsm_gloo():
  """world_size 2 on CPU (gloo): shard the start states by rows, sweep independently, one all_gather"""
  import torch.multiprocessing as mp
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_gloo_fbsm_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  got = q.get(timeout=300)
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  want = _optimizer("CANCERTREATMENT", 60).host_solve_batch(_five_starts())
  for k in ("x", "u", "adj", "iters", "status"):
    assert got[k].shape == want[k].shape, k
    assert np.array_equal(got[k], want[k]), k


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", SYSTEMS)
def test_gpu_fbsm_matches_reference_fixture(name):
  g = _gold(name)
  opt = _optimizer(name, int(g["N"]))
  r = opt.solve_batch(g["x0"])
  torch.cuda.synchronize()
  assert (r["status"] == 0).all()
  assert r["iters"].cpu().tolist() == g["sweeps"].tolist()
  for k in ("x", "u", "adj"):
    _close(r[k].cpu().numpy(), g[k], f"{name} {k}")
  if np.array_equal(g["x0"][0], np.asarray(opt.system.x_0, dtype=np.float64)):
    s = opt.solve()  # the reference's call: one start state, numpy result
    for k in ("x", "u", "adj"):
      assert s[k].shape == g[k][0].shape
      _close(s[k], g[k][0], f"{name} solve() {k}")


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,B", [("CANCERTREATMENT", 1000, 4096), ("HIVTREATMENT", 200, 1000), ("BEARPOPULATIONS", 100, 777),
                                      ("PREDATORPREY", 100, 65), ("INVASIVEPLANT", 10, 333)])
def test_gpu_fbsm_random_start_states_against_oracle(name, N, B):
  """ragged batch sizes; a sample of rows is re-solved by the NumPy oracle; whole-batch properties for the rest"""
  opt = _optimizer(name, N)
  rng = np.random.Generator(np.random.PCG64(7))
  x0d = np.asarray(opt.system.x_0, dtype=np.float64)
  spread = 0.02 if name == "PREDATORPREY" else 0.1  # the secant start values are tuned to the default start state
  x0 = x0d * (1.0 + spread * rng.uniform(-1, 1, size=(B, x0d.shape[0])))
  r = opt.solve_batch(x0)
  torch.cuda.synchronize()
  x, u, adj = (r[k].cpu().numpy() for k in ("x", "u", "adj"))
  assert (r["status"].cpu().numpy() == 0).all()
  assert np.array_equal(x[:, 0, :], x0)                                   # start state is kept exactly
  adj_T = np.zeros_like(x0d) if opt.adj_T is None else opt.adj_T
  free = [k for k in range(x0d.shape[0]) if k != opt.term_cdtion_state]
  assert np.array_equal(adj[:, -1, free], np.broadcast_to(adj_T[free], (B, len(free))))  # transversality condition
  assert (u >= opt.char_lb - 1e-15).all() and (u <= opt.char_ub + 1e-15).all()
  if opt.terminal_cdtion:
    assert np.abs(x[:, -1, opt.term_cdtion_state] - opt.term_value).max() <= 1e-10
  iters = r["iters"].cpu().numpy()
  for row in (0, B // 2, B - 1):
    s = OF.solve(name, N, x0[row])
    assert s["sweeps"] == int(iters[row])
    for k, got in (("x", x), ("u", u), ("adj", adj)):
      _close(got[row], s[k], f"{name} row {row} {k}")


@pytest.mark.gpu
def test_gpu_fbsm_empty_batch():
  opt = _optimizer("SIMPLECASE", 10)
  r = opt.solve_batch(np.zeros((0, 1)))
  assert r["x"].shape == (0, 11, 1) and r["u"].shape == (0, 11, 1)
