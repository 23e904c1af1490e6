import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
for optid, name in ((PR.TRAPEZOIDAL, "trap"), (PR.HERMITE_SIMPSON, "hs")):
  tr = PR.Transcription(SystemType.CARTPOLE(), optid, "HEUN", 100, 1)
  eng = Engine(tr.desc())
  for B in (1, 64, 1024, 8192):
    x0 = PR.sample_x0(tr.system, B, device="cuda")
    z0, lb, ub = PR.build_batch(tr, x0)
    out = eng.ipm_solve(z0, lb, ub); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.ipm_solve(z0, lb, ub, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    print(f"{name} B={B}: {ms:.2f} ms -> {B/ms*1e3:.0f} solves/s; status {dict(zip(*np.unique(st, return_counts=True)))} iters min/med/max {it.min()}/{np.median(it)}/{it.max()} obj0 {float(out['obj'][0]):.10f}", flush=True)
  # K1 alone
  for B in (1024, 8192, 32768):
    z = torch.randn(B, tr.nvars, dtype=torch.float64, device="cuda") * 0.1
    lam = torch.randn(B, tr.ncon, dtype=torch.float64, device="cuda")
    r = eng.eval(z); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.eval(z, out=r)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    s = eng.sizes
    bytes_ = B * 8 * (2 * s.nvars + s.ncon + s.jac_block_doubles + 1)
    print(f"{name} K1 B={B}: {ms*1e3:.1f} us, {bytes_/ms/1e6:.0f} GB/s algorithmic", flush=True)
