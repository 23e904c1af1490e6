"""Entry point: mirror of the reference's run.py:20-53.

    python run.py --system=CARTPOLE --optimizer=COLLOCATION --intervals=100 [--batch=1024]
"""
import random

import numpy as np

from myriad_b200.useful_scripts import run_setup, run_trajectory_opt


def main():
  hp, cfg = run_setup()
  random.seed(hp.seed)
  np.random.seed(hp.seed)
  cost, defect = run_trajectory_opt(hp, cfg, save_as='traj_opt_example.pdf')
  if np.ndim(cost) == 0:
    print("cost", cost, "defect", defect)
  else:
    print("costs[:8]", np.asarray(cost)[:8])


if __name__ == '__main__':
  main()
