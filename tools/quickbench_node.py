"""C5: CARTPOLE with neural-ODE MLP dynamics (3x64), trapezoidal collocation N=100: K1 (tensor-core MLP pass) and batched solves."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from tests.cases import product_system
quads = sys.argv[1].split(",") if len(sys.argv) > 1 else ["trap"]
Bs = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 64, 1024, 4096]
for name in quads:
  optid = PR.TRAPEZOIDAL if name == "trap" else PR.HERMITE_SIMPSON
  tr = PR.Transcription(product_system("NODE_CARTPOLE"), optid, "HEUN", 100, 1)
  eng = Engine(tr.desc())
  s = eng.sizes
  # K1 alone: flops of the MLP pass = 2 * (sum of layer products) * (1 value + NW tangent columns), + reverse pass with Hessian
  for B in (1024, 4096):
    z = torch.randn(B, tr.nvars, dtype=torch.float64, device="cuda") * 0.1
    lam = torch.randn(B, tr.ncon, dtype=torch.float64, device="cuda")
    for hess in (False, True):
      r = eng.eval(z, lam if hess else None, hessian=hess); torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(5): eng.eval(z, lam if hess else None, hessian=hess, out=r)
      e1.record(); torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / 5
      H = tr.system.hidden
      sizes = [tr.n + tr.m] + H + [tr.n]
      mac = sum(a * b for a, b in zip(sizes[:-1], sizes[1:]))
      mac_inner = sum(a * b for a, b in zip(sizes[1:-1], sizes[2:]))  # layers past the first (tangents of layer 1 are the weights)
      cols = 1 + (tr.n + tr.m)
      flops = 2 * (mac + (tr.n + tr.m) * mac_inner + (mac_inner if hess else 0)) * s.nodes * B
      print(f"{name} NODE K1 hess={hess} B={B}: {ms*1e3:.1f} us, {flops/ms/1e9:.2f} TFLOP/s (fp64, MLP products only)", flush=True)
  for B in Bs:
    x0 = PR.sample_x0(tr.system, B, device="cuda")
    z0, lb, ub = PR.build_batch(tr, x0)
    out = eng.ipm_solve(z0, lb, ub); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.ipm_solve(z0, lb, ub, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    print(f"{name} NODE B={B}: {ms:.2f} ms -> {B/ms*1e3:.0f} solves/s; status {dict(zip(*np.unique(st, return_counts=True)))} iters min/med/max {it.min()}/{np.median(it)}/{it.max()} obj0 {float(out['obj'][0]):.10f} cinf {float(out['con_inf'].max()):.1e}", flush=True)
