"""ncu driver for the non-headline configs: python tools/profile_other.py c3|hs [B]"""
import sys, torch
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
what = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 592
if what == "c3":
  tr = PR.Transcription(SystemType.VANDERPOL(), PR.SHOOTING, "HEUN", 1, 50)
else:
  tr = PR.Transcription(SystemType.CARTPOLE(), PR.HERMITE_SIMPSON, "RK4", 100, 1)
eng = Engine(tr.desc())
x0 = PR.sample_x0(tr.system, B, device="cuda")
z0, lb, ub = PR.build_batch(tr, x0)
out = eng.ipm_solve(z0, lb, ub)
torch.cuda.synchronize()
print("ok", int((out["status"] == 0).sum()), "/", B)
