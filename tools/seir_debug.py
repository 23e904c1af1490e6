"""Debug aid: one solved-fixture case through the CUDA IPM (library selected by MYR_LIB), prints status / iterations / errors."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from tests.cases import load, product_transcription
from myriad_b200.engine import Engine
for case in sys.argv[1:]:
  fx = load(case); tr = product_transcription(case); eng = Engine(tr.desc())
  d = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
  out = eng.ipm_solve(d(fx["guess"][None]), d(fx["bounds"][None, :, 0]), d(fx["bounds"][None, :, 1]))
  torch.cuda.synchronize()
  print(case, "status", int(out["status"][0]), "iters", int(out["iters"][0]), "obj %.12f" % float(out["obj"][0]), "kkt %.3e" % float(out["kkt_err"][0]),
        "cinf %.2e" % float(out["con_inf"][0]), "ref", fx.get("sol_cost"), flush=True)
