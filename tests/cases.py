"""Shared list of golden cases (mirrors oracle/make_golden.py:CASES without importing the
reference): name -> (system, optimizer, quadrature, integration_method, intervals, cpi)."""
import os

import numpy as np

from oracle.make_golden import CASES, SOLVE_CASES  # noqa: F401  (pure data; importing it does not touch /root/reference)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(case):
  return dict(np.load(os.path.join(GOLDEN, case + ".npz")))
