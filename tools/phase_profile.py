import os, sys, ctypes as C, torch, numpy as np
sys.path.insert(0, '.')
from myriad_b200 import problems as PR, _lib as ML
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
tr = PR.Transcription(SystemType.CARTPOLE(), PR.TRAPEZOIDAL, "HEUN", 100, 1)
eng = Engine(tr.desc())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x0 = PR.sample_x0(tr.system, B, device="cuda")
z0, lb, ub = PR.build_batch(tr, x0)
out = eng.ipm_solve(z0, lb, ub); torch.cuda.synchronize()
buf = (C.c_double * 16)()
ML.lib().myr_debug_phase_cycles(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = eng.ipm_solve(z0, lb, ub, out=out); e1.record(); torch.cuda.synchronize()
ML.lib().myr_debug_phase_cycles(buf, 0)
v = np.array(list(buf)); it = float(out["iters"].sum())
names = ["K1 eval", "constraints+dual residual+fused reduce", "sigma/rb", "KKT total (incl. inertia retries)", "step sizes", "line search", "  kkt: node inverse+role products", "  kkt: stage assembly", "  kkt: block CR", "accept+loop"]
tot = v[[0, 1, 2, 3, 4, 5, 9]].sum()
print(f"B={B}: {e0.elapsed_time(e1):.2f} ms, iterations {it:.0f}, cycles/iteration (thread 0 of each CTA) {tot/it:.0f}")
for n, c in zip(names, v[:10]): print(f"  {n:32s} {c/it:9.0f} cyc/iter  {100*c/tot:5.1f}%")
for n, c in zip(["CR level0 elim", "CR level0 update", "CR level1 elim", "CR level1 update", "CR levels>=2 elim", "CR levels>=2 update"], v[10:16]):
  print(f"  {n:32s} {c/it:9.0f} cyc/iter")
