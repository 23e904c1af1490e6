"""Shared list of golden cases (mirrors oracle/make_golden.py:CASES without importing the
reference): name -> (system, optimizer, quadrature, integration_method, intervals, cpi)."""
import os

import numpy as np

from oracle.make_golden import CASES, SOLVE_CASES  # noqa: F401  (pure data; importing it does not touch /root/reference)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(case):
  return dict(np.load(os.path.join(GOLDEN, case + ".npz")))


def product_system(sysname):
  """myriad_b200 system object for a case's system name (NODE_<name>: NodeSystem around the true system with the
  committed fixture weights)."""
  from myriad_b200.systems import NodeSystem, SystemType
  if sysname.startswith("NODE_"):
    from oracle.systems import golden_node_weights
    params = {}
    for i, (w, b) in enumerate(golden_node_weights(sysname)):
      params["linear" if i == 0 else f"linear_{i}"] = {"w": w, "b": b}
    return NodeSystem(params, SystemType[sysname[5:]]())
  return SystemType[sysname]()


def product_transcription(case):
  from myriad_b200 import problems as PR
  sysname, opt, quad, meth, intervals, cpi = CASES[case]
  optid = PR.SHOOTING if opt == "SHOOTING" else (PR.TRAPEZOIDAL if quad == "TRAPEZOIDAL" else PR.HERMITE_SIMPSON)
  return PR.Transcription(product_system(sysname), optid, meth, intervals, cpi)
