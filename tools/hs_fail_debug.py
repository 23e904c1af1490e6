"""Debug aid: statuses of the instances that do not reach ST_SOLVED in a batched solve (library via MYR_LIB)."""
import os, sys, torch, numpy as np
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
quad = sys.argv[1] if len(sys.argv) > 1 else "hs"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
tr = PR.Transcription(SystemType.CARTPOLE(), PR.TRAPEZOIDAL if quad == "trap" else PR.HERMITE_SIMPSON, "HEUN", 100, 1)
eng = Engine(tr.desc())
x0 = PR.sample_x0(tr.system, B, device="cuda")
z0, lb, ub = PR.build_batch(tr, x0)
out = eng.ipm_solve(z0, lb, ub); torch.cuda.synchronize()
st = out["status"].cpu().numpy(); bad = np.where(st != 0)[0]
print("not solved:", bad, st[bad], "iters", out["iters"].cpu().numpy()[bad], "kkt", out["kkt_err"].cpu().numpy()[bad], "cinf", out["con_inf"].cpu().numpy()[bad], "obj", out["obj"].cpu().numpy()[bad])
it = out["iters"].cpu().numpy(); print("iters pct", np.percentile(it, [50, 99, 99.9, 100]))
