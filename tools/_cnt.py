import os, sys, torch, numpy as np
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
tr = PR.Transcription(SystemType.CARTPOLE(), PR.TRAPEZOIDAL, "HEUN", 100, 1)
eng = Engine(tr.desc())
x0 = PR.sample_x0(tr.system, 1024, device="cuda")
z0, lb, ub = PR.build_batch(tr, x0)
out = eng.ipm_solve(z0, lb, ub); torch.cuda.synchronize()
it = out["iters"].cpu().numpy()
k = it // 10000; i = it % 10000
print("iters sum", i.sum(), "kkt solves sum", k.sum(), "ratio", k.sum() / i.sum(), "max ratio", (k / np.maximum(i, 1)).max())
