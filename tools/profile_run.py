"""Tiny driver for ncu captures: python tools/profile_run.py [trap|hs] [B] [ipm|k1]"""
import sys, torch
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
quad = sys.argv[1] if len(sys.argv) > 1 else "trap"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
what = sys.argv[3] if len(sys.argv) > 3 else "ipm"
tr = PR.Transcription(SystemType.CARTPOLE(), PR.TRAPEZOIDAL if quad == "trap" else PR.HERMITE_SIMPSON, "HEUN", 100, 1)
eng = Engine(tr.desc())
x0 = PR.sample_x0(tr.system, B, device="cuda")
z0, lb, ub = PR.build_batch(tr, x0)
if what == "ipm":
  out = eng.ipm_solve(z0, lb, ub)
  torch.cuda.synchronize()
  print("status ok", int((out["status"] == 0).sum()), "/", B)
else:
  r = eng.eval(z0)
  for _ in range(3):
    eng.eval(z0, out=r)
  torch.cuda.synchronize()
