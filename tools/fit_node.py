#!/usr/bin/env python
"""Fit the reference's NODE architecture (myriad/neural_ode/create_node.py:110-117: Linear -> sigmoid per hidden layer,
then Linear; no input normalisation) to a system's true dynamics by regression on f(x, u), so that BASELINE config C5
(CARTPOLE with neural_ode MLP dynamics 3x64) has weights for which the planning problem is well posed.  The reference
trains the same network on trajectory data (node_training.py); there is no dataset / checkpoint here (no network), so
the fit is on samples of the vector field itself.  Output: tests/golden/node_<system>_<hidden>.npz with haiku-style keys
linear/w, linear/b, linear_1/w, ... (create_node.py:124-131).

    python tools/fit_node.py CARTPOLE 64,64,64 [seconds]
"""
import sys, time
import numpy as np
import torch

sys.path.insert(0, ".")
from oracle.systems import make_system, haiku_style_mlp_weights

name = sys.argv[1] if len(sys.argv) > 1 else "CARTPOLE"
hidden = [int(s) for s in (sys.argv[2] if len(sys.argv) > 2 else "64,64,64").split(",")]
budget = float(sys.argv[3]) if len(sys.argv) > 3 else 600.0
torch.manual_seed(0)
torch.set_default_dtype(torch.float64)
sysm = make_system(name)
n, m = sysm.n, sysm.m
b = np.asarray(sysm.bounds, dtype=np.float64)
lo, hi = torch.as_tensor(b[:, 0]), torch.as_tensor(b[:, 1])
lo = torch.where(torch.isfinite(lo), lo, torch.full_like(lo, -5.0))
hi = torch.where(torch.isfinite(hi), hi, torch.full_like(hi, 5.0))


def sample(N):
  v = lo + (hi - lo) * torch.rand(N, n + m)
  return v, sysm.dynamics(v[:, :n], v[:, n:])


W = [(torch.tensor(w, requires_grad=True), torch.tensor(bb, requires_grad=True))
     for w, bb in haiku_style_mlp_weights(n + m, hidden, n, seed=42)]
params = [p for wb in W for p in wb]


def net(v):
  h = v
  for i, (w, bb) in enumerate(W):
    h = h @ w + bb
    if i + 1 < len(W):
      h = torch.sigmoid(h)
  return h


Xtr, Ytr = sample(32768)
Xte, Yte = sample(8192)
scale = Ytr.std(0)
t0 = time.time()
opt = torch.optim.Adam(params, lr=3e-3)
it = 0
while time.time() - t0 < 0.5 * budget:
  idx = torch.randint(0, Xtr.shape[0], (2048,))
  loss = (((net(Xtr[idx]) - Ytr[idx]) / scale) ** 2).mean()
  opt.zero_grad(); loss.backward(); opt.step()
  it += 1
  if it % 500 == 0:
    print(f"adam {it} loss {float(loss):.3e}", flush=True)
opt = torch.optim.LBFGS(params, lr=1.0, max_iter=20, history_size=50, line_search_fn="strong_wolfe")
while time.time() - t0 < budget:
  def closure():
    opt.zero_grad()
    l = (((net(Xtr) - Ytr) / scale) ** 2).mean()
    l.backward()
    return l
  l = opt.step(closure)
  with torch.no_grad():
    te = (((net(Xte) - Yte) / scale) ** 2).mean()
  print(f"lbfgs train {float(l):.3e} test {float(te):.3e}", flush=True)
out = {}
for i, (w, bb) in enumerate(W):
  key = "linear" if i == 0 else f"linear_{i}"
  out[key + "/w"] = w.detach().numpy()
  out[key + "/b"] = bb.detach().numpy()
fn = f"tests/golden/node_{name.lower()}_{'x'.join(str(h) for h in hidden)}.npz"
np.savez_compressed(fn, **out)
print("saved", fn)
