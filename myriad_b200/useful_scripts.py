"""Entry helpers: mirror of myriad/useful_scripts.py:26-76 (run_trajectory_opt) and :104-137 (run_setup).

``run_setup`` re-implements the simple_parsing behaviour the reference relies on with argparse: one flag per
dataclass field, enums parsed by member name (``--system=CARTPOLE --optimizer=COLLOCATION``)."""
from __future__ import annotations

import argparse
import dataclasses
import enum
import pickle as pkl
import typing
from typing import Optional, Tuple

import numpy as np

from myriad_b200.config import Config, HParams
from myriad_b200.trajectory_optimizers import get_optimizer
from myriad_b200.utils import get_defect, get_state_trajectory_and_cost


def _add_dataclass_args(parser: argparse.ArgumentParser, cls) -> None:
  hints = typing.get_type_hints(cls)
  for f in dataclasses.fields(cls):
    tp = hints[f.name]
    name = "--" + f.name
    if isinstance(tp, type) and issubclass(tp, enum.Enum):
      parser.add_argument(name, type=lambda s, tp=tp: tp[s.split(".")[-1]], default=f.default,
                          help=f"one of {[m.name for m in tp]}")
    elif tp is bool:
      parser.add_argument(name, type=lambda s: str(s).lower() in ("1", "true", "yes", "y"), default=f.default,
                          nargs="?", const=True)
    elif tp in (int, float, str):
      parser.add_argument(name, type=tp, default=f.default)
    else:  # Tuple[...] fields
      parser.add_argument(name, type=float, nargs="+", default=f.default)


def run_setup(argv=None) -> Tuple[HParams, Config]:
  parser = argparse.ArgumentParser()
  _add_dataclass_args(parser, HParams)
  _add_dataclass_args(parser, Config)
  args = vars(parser.parse_args(argv))
  hp_fields = {f.name for f in dataclasses.fields(HParams)}
  cfg_fields = {f.name for f in dataclasses.fields(Config)}
  for k in ("hidden_layers", "figsize"):
    if isinstance(args[k], list):
      args[k] = tuple(int(v) if k == "hidden_layers" else v for v in args[k])
  hp = HParams(**{k: v for k, v in args.items() if k in hp_fields})
  cfg = Config(**{k: v for k, v in args.items() if k in cfg_fields})
  print(hp)
  print(cfg)
  np.random.seed(hp.seed)
  return hp, cfg


def run_trajectory_opt(hp: HParams, cfg: Config, save_as: Optional[str] = None, params_path: Optional[str] = None):
  """Solve, then re-integrate the TRUE system under the solved controls; returns (cost, defect) like the reference.
  With hp.batch > 1 the solve is batched over perturbed start states and per-instance arrays are returned."""
  if params_path is not None:
    params = pkl.load(open(params_path, 'rb'))
    system = hp.system(**params)
    print("loaded params:", params)
  else:
    system = hp.system()
    print("made default system")
  optimizer = get_optimizer(hp, cfg, system)
  true_system = hp.system()
  if hp.batch > 1:
    import torch
    from myriad_b200 import problems as PR
    from myriad_b200.utils import get_state_trajectory_and_cost_batch
    x0s = PR.sample_x0(true_system, hp.batch, seed=hp.seed, spread=hp.start_spread, device="cuda")
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    sol = optimizer.solve_batch(x0s)
    xs, costs = get_state_trajectory_and_cost_batch(hp, true_system, x0s, sol['u'])
    t1.record(); torch.cuda.synchronize()
    ok = (sol['status'] == 0)
    if cfg.verbose:
      print(f"solved {int(ok.sum())}/{hp.batch} instances in {t0.elapsed_time(t1):.2f} ms "
            f"({hp.batch / t0.elapsed_time(t1) * 1e3:.0f} solves/s), iterations median {int(sol['iters'].median())}")
    defects = None
    if true_system.x_T is not None:
      idx = [i for i, v in enumerate(true_system.x_T) if v is not None]
      tgt = torch.as_tensor([float(true_system.x_T[i]) for i in idx], dtype=torch.float64, device=xs.device)
      defects = (xs[:, -1, idx] - tgt).cpu().numpy()
    return costs.cpu().numpy(), defects
  solution = optimizer.solve()
  u = solution['u']
  opt_x, c = get_state_trajectory_and_cost(hp, true_system, true_system.x_0, u)
  defect = get_defect(true_system, opt_x)
  if cfg.plot:
    try:
      from myriad_b200.plotting import plot
      plot(hp, true_system, data={'x': opt_x, 'u': u, 'cost': c, 'defect': defect}, save_as=save_as)
    except ImportError:
      pass  # matplotlib is optional (not in this image); plotting is presentation only
  return c, defect
