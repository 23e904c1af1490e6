"""BASELINE config C4: CANCERTREATMENT shooting 1x100, max_iter=500, batch sweep 2^6 .. 2^18 on one GPU."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
tr = PR.Transcription(SystemType.CANCERTREATMENT(), PR.SHOOTING, "HEUN", 1, 100)
eng = Engine(tr.desc())
print("workspace doubles/instance", eng.sizes.ipm_workspace_doubles)
for e in range(6, 19, 2):
  B = 1 << e
  x0 = PR.sample_x0(tr.system, B, device="cuda")
  z0, lb, ub = PR.build_batch(tr, x0)
  out = eng.ipm_solve(z0, lb, ub, max_iter=500); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); out = eng.ipm_solve(z0, lb, ub, max_iter=500, out=out); e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  st = out["status"].cpu().numpy()
  print(f"C4 B=2^{e}={B}: {ms:.2f} ms -> {B/ms*1e3:.0f} solves/s; solved {int((st==0).sum())}/{B}; cinf max {float(out['con_inf'].max()):.1e}; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
