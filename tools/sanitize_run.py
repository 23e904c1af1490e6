"""Small invocations of every kernel for compute-sanitizer (racecheck / memcheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import sys, torch
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
from tests.cases import product_system

def run(system, optid, meth, N, cpi, B, label):
  tr = PR.Transcription(system, optid, meth, N, cpi)
  eng = Engine(tr.desc())
  x0 = PR.sample_x0(tr.system, B, device="cuda")
  z0, lb, ub = PR.build_batch(tr, x0)
  lam = torch.randn(B, tr.ncon, dtype=torch.float64, device="cuda")
  if optid != PR.SHOOTING:
    r = eng.eval(z0, lam, hessian=True)
    eng.jtvec(r.Jblk, lam)
  else:
    r = eng.eval(z0)
    eng.jtvec(r.Jblk, lam)
  out = eng.ipm_solve(z0, lb, ub, max_iter=30)
  _, u = tr.unravel(out["z"])
  eng.rollout_cost(u.contiguous(), x0)
  torch.cuda.synchronize()
  print(label, "status", out["status"].tolist(), "iters", out["iters"].tolist(), flush=True)

cp = SystemType.CARTPOLE()
run(cp, PR.TRAPEZOIDAL, "HEUN", 12, 1, 3, "trap")
run(cp, PR.HERMITE_SIMPSON, "RK4", 6, 1, 2, "hs")
run(cp, PR.SHOOTING, "HEUN", 3, 4, 3, "shooting heun")
run(SystemType.VANDERPOL(), PR.SHOOTING, "RK4", 2, 3, 2, "shooting rk4")
run(product_system("NODE_CARTPOLE"), PR.TRAPEZOIDAL, "HEUN", 9, 1, 2, "node trap")
# K1 with a subset of outputs (f and c only) on the NODE path, and the forward-backward sweep kernel (incl. the secant loop
# and the discrete branch)
tr = PR.Transcription(product_system("NODE_CARTPOLE"), PR.HERMITE_SIMPSON, "RK4", 6, 1)
eng = Engine(tr.desc())
z0, _, _ = PR.build_batch(tr, PR.sample_x0(tr.system, 2, device="cuda"))
eng.eval(z0, jac=False)
torch.cuda.synchronize()
from myriad_b200.config import Config, HParams, OptimizerType
from myriad_b200.trajectory_optimizers import get_optimizer
for name, N in (("CANCERTREATMENT", 40), ("PREDATORPREY", 30), ("BEARPOPULATIONS", 20), ("INVASIVEPLANT", 10)):
  hp = HParams(system=SystemType[name], optimizer=OptimizerType.FBSM, fbsm_intervals=N)
  opt = get_optimizer(hp, Config(verbose=False, plot=False), hp.system())
  x0 = torch.as_tensor(opt.system.x_0).reshape(1, -1).repeat(37, 1)
  r = opt.solve_batch(x0)
  torch.cuda.synchronize()
  print("fbsm", name, "sweeps", int(r["iters"][0]), "status", int(r["status"].abs().max()), flush=True)
print("done")
