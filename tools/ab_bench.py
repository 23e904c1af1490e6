"""A/B timing of the headline solve (CARTPOLE trapezoid N=100) for the library selected by MYR_LIB."""
import os, sys, torch, numpy as np
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
tag = os.environ.get("MYR_LIB", "default")
quads = sys.argv[1].split(",") if len(sys.argv) > 1 else ["trap"]
for name in quads:
  tr = PR.Transcription(SystemType.CARTPOLE(), PR.TRAPEZOIDAL if name == "trap" else PR.HERMITE_SIMPSON, "HEUN", 100, 1)
  eng = Engine(tr.desc())
  for B in (1024, 4096, 8192):
    off = int(os.environ.get("ROW_OFFSET", "0"))   # rows [off, off + B) of the seeded draw (another rank's shard)
    x0 = PR.sample_x0(tr.system, off + B, device="cuda")[off:].contiguous()
    z0, lb, ub = PR.build_batch(tr, x0)
    out = eng.ipm_solve(z0, lb, ub); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); out = eng.ipm_solve(z0, lb, ub, out=out); e1.record(); torch.cuda.synchronize()
      best = min(best, e0.elapsed_time(e1))
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    print(f"[{os.path.basename(tag)}] {name} B={B}: {best:.2f} ms -> {B/best*1e3:.0f} solves/s; ok {int((st==0).sum())}/{B} iters med {np.median(it)} sum {it.sum()} obj0 {float(out['obj'][0]):.10f}", flush=True)
