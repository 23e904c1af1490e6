import sys, torch
sys.path.insert(0, '.')
from myriad_b200 import problems as PR
from myriad_b200.engine import Engine
from myriad_b200.systems import SystemType
for optid, name in ((PR.TRAPEZOIDAL, "trap"), (PR.HERMITE_SIMPSON, "hs")):
  tr = PR.Transcription(SystemType.CARTPOLE(), optid, "HEUN", 100, 1)
  eng = Engine(tr.desc())
  for B in (1024, 8192, 32768):
    z = torch.randn(B, tr.nvars, dtype=torch.float64, device="cuda") * 0.1
    r = eng.eval(z); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.eval(z, out=r)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    s = eng.sizes
    bytes_ = B * 8 * (2 * s.nvars + s.ncon + s.jac_block_doubles + 1)
    print(f"{name} K1 B={B}: {ms*1e3:.1f} us, {bytes_/ms/1e6:.0f} GB/s algorithmic ({bytes_/ms/1e6/6536.7*100:.1f}% of measured HBM peak)", flush=True)
