"""ncu driver for the NODE (neural-ODE MLP, fp64 tensor-core) K1: python tools/profile_node.py [B]"""
import sys, torch
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from tests.cases import product_system
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
tr = PR.Transcription(product_system("NODE_CARTPOLE"), PR.TRAPEZOIDAL, "HEUN", 100, 1)
eng = Engine(tr.desc())
z = torch.randn(B, tr.nvars, dtype=torch.float64, device="cuda") * 0.1
lam = torch.randn(B, tr.ncon, dtype=torch.float64, device="cuda")
r = eng.eval(z, lam, hessian=True)
for _ in range(2):
  eng.eval(z, lam, hessian=True, out=r)
torch.cuda.synchronize()
