import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
def run(name, system, K, cpi, meth, Bs):
  tr = PR.Transcription(system, PR.SHOOTING, meth, K, cpi)
  eng = Engine(tr.desc())
  for B in Bs:
    x0 = PR.sample_x0(tr.system, B, device="cuda")
    z0, lb, ub = PR.build_batch(tr, x0)
    out = eng.ipm_solve(z0, lb, ub); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.ipm_solve(z0, lb, ub, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    print(f"{name} B={B}: {ms:.2f} ms -> {B/ms*1e3:.0f} solves/s; status {dict(zip(*np.unique(st, return_counts=True)))} iters {it.min()}/{np.median(it)}/{it.max()} obj0 {float(out['obj'][0]):.10f} cinf max {float(out['con_inf'].max()):.1e}", flush=True)
run("C3 VANDERPOL 1x50 HEUN", SystemType.VANDERPOL(), 1, 50, "HEUN", (1, 1024, 8192))
run("C4 CANCER 1x100 HEUN", SystemType.CANCERTREATMENT(), 1, 100, "HEUN", (64, 1024, 16384, 65536))
run("C1 SIMPLECASE 10x100 HEUN", SystemType.SIMPLECASE(), 10, 100, "HEUN", (1, 1024))
