"""CPU tests of the host-side mirror of the reference's plugin surface (no GPU): enums, HParams derived fields,
run_setup flag parsing, factory error behaviour, sharding helpers (world_size-2 gloo)."""
import os
import sys

import numpy as np
import pytest
import torch


def test_hparams_post_init_semantics():
  from myriad_b200.config import HParams, IntegrationMethod, NLPSolverType, OptimizerType
  from myriad_b200.systems import SystemType
  hp = HParams(system=SystemType.CARTPOLE, optimizer=OptimizerType.COLLOCATION, intervals=100, controls_per_interval=7)
  assert hp.controls_per_interval == 1  # myriad/config.py:98-99
  assert hp.num_steps == 100 and abs(hp.stepsize - 0.02) < 1e-15 and hp.state_size == 4 and hp.control_size == 1
  hp = HParams(nlpsolver=NLPSolverType.EXTRAGRADIENT, max_iter=50)
  assert hp.max_iter == 500  # :100-101
  hp = HParams()
  assert hp.system == SystemType.CANCERTREATMENT and hp.optimizer == OptimizerType.SHOOTING and hp.controls_per_interval == 100
  assert hp.integration_method == IntegrationMethod.HEUN and hp.minibatch_size == 3  # :112
  assert IntegrationMethod.EULER.value == "CONSTANT" and IntegrationMethod.HEUN.value == "LINEAR"


def test_run_setup_flags_like_the_reference():
  from myriad_b200.config import OptimizerType, QuadratureRule
  from myriad_b200.systems import SystemType
  from myriad_b200.useful_scripts import run_setup
  hp, cfg = run_setup(["--system=CARTPOLE", "--optimizer=COLLOCATION", "--intervals=100", "--quadrature_rule=HERMITE_SIMPSON",
                       "--max_iter=500", "--verbose=false", "--plot=false", "--batch=64"])
  assert hp.system == SystemType.CARTPOLE and hp.optimizer == OptimizerType.COLLOCATION and hp.intervals == 100
  assert hp.quadrature_rule == QuadratureRule.HERMITE_SIMPSON and hp.max_iter == 500 and hp.batch == 64
  assert cfg.verbose is False and cfg.plot is False


def test_system_descriptors_match_reference_tables():
  from myriad_b200.systems import SystemType
  from tests.cases import load
  cp = SystemType.CARTPOLE()
  assert cp.T == 2.0 and np.allclose(cp.x_T, [1.0, np.pi, 0, 0]) and cp.bounds.shape == (5, 2)
  fx = load("c2_cartpole_trap_100")
  assert np.array_equal(fx["bounds"][4:8, 0], cp.bounds[:4, 0])  # node 1 carries the plain state bounds
  ct = SystemType.CANCERTREATMENT()
  assert ct.x_T is None and ct.T == 20 and np.allclose(ct.x_0, [0.975])
  rl = SystemType.ROCKETLANDING()
  assert rl.state_size == 6 and rl.control_size == 2 and rl.T == 16.0
  ip = SystemType.INVASIVEPLANT()   # discrete system: FBSM only; the direct optimizers reject it like the reference (base.py:66-67)
  assert ip.discrete and ip.state_size == 5 and ip.control_size == 5 and ip.T == 10.0


def test_shard_ranges_cover_everything():
  from myriad_b200.distributed import shard_range
  for total in (1, 7, 64, 1000):
    for w in (1, 2, 3, 8):
      spans = [shard_range(total, r, w) for r in range(w)]
      assert spans[0][0] == 0 and spans[-1][1] == total
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _gloo_worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  import ctypes as C
  import torch.distributed as dist
  from myriad_b200 import _lib as ML
  from myriad_b200 import distributed as D
  dist.init_process_group("gloo", rank=rank, world_size=world)
  # shard 5 CARTPOLE N=10 instances over 2 ranks (UNEQUAL shards: 3 + 2); solve each shard with the HOST twin; gather once
  from tests.test_cpu_abi_and_twin import _host_ipm
  d = ML.make_desc("CARTPOLE", ML.OPT_TRAPEZOIDAL, "HEUN", 10)
  s = ML.problem_sizes(d)
  z0, lb, ub = _tiny_batch(5)
  lo, hi = D.shard_range(5, rank, world)
  out = _host_ipm(ML, d, s, np.ascontiguousarray(z0[lo:hi]), np.ascontiguousarray(lb[lo:hi]), np.ascontiguousarray(ub[lo:hi]))
  t = lambda a: torch.as_tensor(a)
  packed = D.pack_solution(t(out["z"]), t(out["lam"]), t(out["obj"]), t(out["obj"]), t(out["status"]), t(out["iters"]))
  allp = D.gather_solutions(packed)
  if rank == 0:
    q.put(allp.numpy())
  dist.barrier()
  dist.destroy_process_group()


def _tiny_batch(B):
  from tests.cases import load
  fx = load("s_cartpole_trap_10")
  rng = np.random.default_rng(5)
  z0 = np.repeat(fx["guess"][None], B, 0).copy()
  lb = np.repeat(fx["bounds"][None, :, 0], B, 0).copy()
  ub = np.repeat(fx["bounds"][None, :, 1], B, 0).copy()
  x0 = 0.05 * rng.standard_normal((B, 4)); x0[0] = 0
  for b in range(B):
    z0[b, :4] = x0[b]; lb[b, :4] = x0[b]; ub[b, :4] = x0[b]
  return z0, lb, ub


def test_two_rank_shard_and_gather_gloo():
  """world_size 2 on CPU (gloo): the N>1 path = shard by rows, solve independently, one all_gather."""
  import torch.multiprocessing as mp
  from myriad_b200 import _lib as ML
  if not os.path.exists(ML.LIB_PATH):
    from myriad_b200 import build
    build.build(verbose=False)
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
  for pr in procs:
    pr.start()
  got = q.get(timeout=120)
  for pr in procs:
    pr.join(timeout=60)
    assert pr.exitcode == 0
  # single-process reference
  from tests.test_cpu_abi_and_twin import _host_ipm
  d = ML.make_desc("CARTPOLE", ML.OPT_TRAPEZOIDAL, "HEUN", 10)
  s = ML.problem_sizes(d)
  z0, lb, ub = _tiny_batch(5)
  out = _host_ipm(ML, d, s, z0, lb, ub)
  assert got.shape == (5, s.nvars + s.ncon + 4)
  np.testing.assert_array_equal(got[:, :s.nvars], out["z"])
  np.testing.assert_array_equal(got[:, -2], out["status"].astype(np.float64))
  assert (out["status"] == 0).all()


def test_node_system_surface_and_descriptor():
  """NodeSystem (node_system.py:14-42): true system's data, MLP layers parsed from haiku-style parameters in layer order
  (create_node.py:124-131), flattened weight vector and descriptor fields of the C ABI; no GPU involved."""
  from myriad_b200 import _lib as ML, problems as PR
  from myriad_b200.neural_ode import NeuralODE, init_params
  from myriad_b200.config import Config, HParams, OptimizerType
  from myriad_b200.systems import NodeSystem, SystemType, mlp_layers
  true = SystemType.CARTPOLE()
  params = init_params(5, (7, 9), 4, seed=1)
  ns = NodeSystem(params, true)
  assert ns.device_name == "NODE_CARTPOLE" and ns.hidden == [7, 9] and ns.T == true.T and np.array_equal(ns.bounds, true.bounds)
  assert ns.theta.size == 5 * 7 + 7 + 7 * 9 + 9 + 9 * 4 + 4
  np.testing.assert_array_equal(ns.theta[:35], params["linear"]["w"].ravel())
  np.testing.assert_array_equal(ns.theta[-4:], params["linear_2"]["b"])
  flat = {f"{k}/{p}": v for k, d in params.items() for p, v in d.items()}  # .npz style keys load as well
  assert [w.shape for w, _ in mlp_layers(flat)] == [(5, 7), (7, 9), (9, 4)]
  with pytest.raises(ValueError):
    NodeSystem(init_params(4, (8,), 4), true)  # wrong input width
  d = PR.Transcription(ns, PR.TRAPEZOIDAL, "HEUN", 10, 1).desc(device="host")
  assert d.system_id == ML.NODE_BASE + ML.SYSTEM_IDS["CARTPOLE"] and d.node_num_hidden == 2 and list(d.node_hidden[:2]) == [7, 9]
  assert d.theta_doubles == ns.theta.size
  s = ML.problem_sizes(d)
  assert (s.n, s.m, s.nvars, s.ncon) == (4, 1, 55, 40)
  d.theta_doubles += 1  # inconsistent sizes are an argument error, not a crash
  with pytest.raises(KeyError):
    ML.problem_sizes(d)
  hp = HParams(system=SystemType.CARTPOLE, optimizer=OptimizerType.COLLOCATION, intervals=10, hidden_layers=(7, 9))
  node = NeuralODE(hp, Config(verbose=False, plot=False), params=params)
  y = node.apply(params, np.zeros(5))
  assert y.shape == (4,)
