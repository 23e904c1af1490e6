"""ncu driver for the NODE (neural-ODE MLP) interior-point solve, BASELINE config C5: python tools/profile_node_ipm.py [B]"""
import sys, torch
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from tests.cases import product_system
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
tr = PR.Transcription(product_system("NODE_CARTPOLE"), PR.TRAPEZOIDAL, "HEUN", 100, 1)
eng = Engine(tr.desc())
x0 = PR.sample_x0(tr.system, B, device="cuda")
z0, lb, ub = PR.build_batch(tr, x0)
out = eng.ipm_solve(z0, lb, ub)
torch.cuda.synchronize()
print("ok", int((out["status"] == 0).sum()), "/", B)
