"""HBM bandwidth probes with torch kernels: pure write (fill_), pure read (sum), copy (read+write)."""
import torch
n = 1 << 27  # 1 GiB of float64
a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
def t(fn, reps=10):
  fn(); torch.cuda.synchronize()
  best = 1e9
  for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
  return best
w = t(lambda: a.fill_(1.0)); r = t(lambda: a.sum()); c = t(lambda: b.copy_(a))
print(f"pure write (fill_ 1 GiB): {n*8/w/1e6:.0f} GB/s; pure read (sum): {n*8/r/1e6:.0f} GB/s; copy (read+write bytes): {2*n*8/c/1e6:.0f} GB/s")
