#!/usr/bin/env python
"""Headline benchmark: trajopt solves/sec, CARTPOLE trapezoidal collocation N=100, batch of random start states.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = one pass of the hot path over one batch: the batched interior-point solve (myr_ipm_solve) of `batch`
instances per GPU plus the post-solve verification rollout of every instance (myr_rollout_cost; what the
reference's run_trajectory_opt returns).  Work is sharded over ranks by instance (weak scaling, no data-path
collective); one NCCL all_gather of the packed solutions closes each step.

Printed JSON (rank 0, one line) follows the contract in the task statement: value = device-timed whole-job
solves/s with inputs resident in HBM; e2e = same metric through the public Python API from pinned HOST buffers
(H2D of the start states, problem construction, solve, rollout, D2H of the WHOLE result dictionary the reference's
solve() returns -- x, u, xs_and_us, lambda, cost -- plus the re-integrated cost, inside the timed region);
roofline = the step's DOMINANT kernel, ipm_kernel: algorithmic fp64 FLOP/s against a DFMA peak measured live (the kernel
is fp64-latency bound, neither "hbm" nor "tensor": bound is reported as "fp64"), with its DRAM traffic read from the
committed ncu summary; roofline_k1 = the rollout+Jacobian kernel K1 (myr_eval), the HBM-bound kernel the north star names,
timed live over a working set larger than L2; cpu_baseline = the CPU oracle (reference algorithm restated, SciPy SLSQP)
on a bounded sample; cpu_baseline_twin = the SAME interior-point algorithm (host twin of the CUDA templates, OpenMP over
instances) on the same rows: the ratio to it is the hardware factor, the ratio to the SLSQP arm is mostly algorithm;
parity_sample = GPU vs SLSQP objectives on the rows both solved.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SYSTEM, INTERVALS = "CARTPOLE", 100
METRIC = "trajopt solves/sec (CARTPOLE collocation N=100, batched)"


def workload_name(quadrature: str) -> str:
  """config.workload, identical in both arms (the reference arm times a bounded sample of the same workload)"""
  return (f"{SYSTEM} COLLOCATION {quadrature} intervals={INTERVALS}, batch=1024 random x0 (seed 2019, spread 0.1), fp64 "
          "(BASELINE.json configs[1])")


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=40)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--batch", type=int, default=1024, help="instances per GPU per step")
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--quadrature", default="TRAPEZOIDAL", choices=["TRAPEZOIDAL", "HERMITE_SIMPSON"])
  ap.add_argument("--cpu-instances", type=int, default=0, help="bounded sample for the CPU baseline (default: one per core)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  return ap.parse_args()


# ----------------------------------------------------------------------------- CPU arms
def run_cpu_oracle(quadrature: str, instances: int, procs: int) -> dict:
  cmd = [sys.executable, "-m", "oracle.cpu_baseline", "--system", SYSTEM, "--optimizer", "COLLOCATION", "--quadrature", quadrature,
         "--intervals", str(INTERVALS), "--instances", str(instances), "--procs", str(procs)]
  env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
  out = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, check=True).stdout
  return json.loads(out.strip().splitlines()[-1])


def reference_arm(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  # each step: one SLSQP solve per host core, in parallel, on the first `cores` instances of the workload.
  # The step count is bounded so the whole arm ends within a few minutes (~20 s per step on CARTPOLE N=100).
  first = run_cpu_oracle(args.quadrature, cores, cores)
  budget_s = 150.0
  n_more = max(0, min(args.steps - 1, int(budget_s / max(first["seconds"], 1e-3)) - 1))
  vals = [first] + [run_cpu_oracle(args.quadrature, cores, cores) for _ in range(n_more)]
  args.warmup = 0
  secs = sum(v["seconds"] for v in vals)
  n = sum(v["instances"] for v in vals)
  value = n / secs
  line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": len(vals),
          "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(vals), "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
          "config": {"workload": workload_name(args.quadrature), "batch_per_step": cores,
                     "sample": f"{cores} of the workload's instances per step (one SLSQP solve per host core)"},
          "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                           "sample": f"{cores} instances per step, one SciPy-SLSQP solve per core (oracle restatement of the "
                                     "reference transcription; IPOPT/jax are not installable here)"},
          "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
  print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
  def __init__(self, index: int):
    self.samples, self.reasons, self.max_mhz = [], set(), None
    self._stop = threading.Event()
    self._index = index
    self._t = threading.Thread(target=self._run, daemon=True)

  def _run(self):
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    while not self._stop.is_set():
      try:
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self._index)],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
        p = [s.strip() for s in out.split(",")]
        self.samples.append(float(p[0]))
        self.max_mhz = float(p[1])
        for nme, v in zip(names, p[2:]):
          if v.lower().startswith("active"):
            self.reasons.add(nme)
      except Exception:
        pass
      self._stop.wait(0.05)

  def __enter__(self):
    self._t.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    self._t.join(timeout=3)

  def summary(self):
    s = sorted(self.samples)
    return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def ncu_traffic_bytes(summary_path: str, kernel_substr: str):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, parsed from a committed ncu summary
  (tools/ncu_summary.py output under profiles/); None when the file or the kernel is missing."""
  try:
    txt = open(os.path.join(ROOT, summary_path)).read()
  except OSError:
    return None
  unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
  for block in txt.split("\n## ")[1:]:
    if kernel_substr not in block.splitlines()[0]:
      continue
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
      for ln in block.splitlines():
        if ln.startswith(key + " "):
          parts = ln.split()
          tot += float(parts[1]) * unit.get(parts[2], 1.0)
          break
      else:
        return None
    return tot
  return None


IPM_PROFILE = "profiles/r2_ipm_trap_B1024_ncu_full.txt"
K1_PROFILE = "profiles/r1_k1_eval_trap_B8192_ncu_full.txt"


def run_cpu_twin(eng, tr, z0, lb, ub, rows: int) -> dict:
  """The SAME algorithm on the host cores: myr_host_ipm_solve (host twin of the CUDA templates, OpenMP over instances,
  one workspace slot per thread) on the first `rows` instances of the workload."""
  import ctypes as C
  import numpy as np
  from myriad_b200 import _lib as ML
  s = eng.sizes
  rows = int(min(rows, z0.shape[0]))
  h = [np.ascontiguousarray(t[:rows].cpu().numpy()) for t in (z0, lb, ub)]
  out = dict(z=np.zeros((rows, s.nvars)), lam=np.zeros((rows, s.ncon)), zL=np.zeros((rows, s.nvars)), zU=np.zeros((rows, s.nvars)),
             obj=np.zeros(rows), kkt=np.zeros(rows), cinf=np.zeros(rows), status=np.zeros(rows, np.int32), iters=np.zeros(rows, np.int32))
  ws = np.zeros(ML.workspace_doubles(s, rows))
  o = ML.MyrIpmOpts()
  o.max_iter = 1000
  p = lambda a: a.ctypes.data_as(C.c_void_p)
  d = tr.desc(device="host")
  t0 = time.perf_counter()
  ML.check(ML.lib().myr_host_ipm_solve(C.byref(d), C.byref(o), rows, p(h[0]), p(h[1]), p(h[2]), p(out["z"]), p(out["lam"]), p(out["zL"]),
                                       p(out["zU"]), p(out["obj"]), p(out["kkt"]), p(out["cinf"]), p(out["status"]), p(out["iters"]),
                                       p(ws), ws.size))
  secs = time.perf_counter() - t0
  return {"value": rows / secs, "unit": "solves/s", "cores": os.cpu_count() or 1, "kind": "port",
          "sample": f"{rows} instances of the same workload through myr_host_ipm_solve (same interior-point templates compiled for "
                    f"the host, OpenMP over instances; {secs:.2f} s wall)",
          "solved": int((out["status"] == 0).sum()), "obj": out["obj"]}


def ipm_flops_per_iteration(sz) -> dict:
  """Algorithmic fp64 flops of one interior-point iteration of one instance: node evaluation with
  Hessian, one KKT solve (node-block LDL^T inverses, Schur complement, block cyclic reduction, back-substitution) and
  one values-only line-search evaluation, for Q nodes of NW variables and St stages of NC rows, k nodes per stage."""
  Q, St, NW, NC, k = sz.nodes, sz.stages, sz.nw, sz.nc, sz.stage_nodes
  node_inv = Q * 2 * NW ** 3
  schur = St * (k * (2 * NC * NW * NW + 2 * NC * NC * NW + 2 * NC * NW) + 2 * NC * NC * NW)
  cr = St * (2 * NC ** 3 + 2 * 2 * NC ** 3 + 2 * NC * NC + 2 * (2 * 2 * NC ** 3 + 2 * NC * NC) + 2 * 2 * NC * NC)
  dz = Q * (2 * 2 * NC * NW + 2 * NW * NW)
  k1 = Q * 400 + Q * 100  # generated f/J/Hessian code of CARTPOLE (~400 flops incl. sincos) + values-only trial evaluation
  vec = 30 * Q * NW      # residuals, barrier terms, fraction-to-boundary, updates
  return {"node_inverse": node_inv, "schur": schur, "block_cr": cr, "dz": dz, "k1": k1, "vector": vec,
          "total": node_inv + schur + cr + dz + k1 + vec}


# ----------------------------------------------------------------------------- B200 arm
def b200_arm(args):
  import torch
  import torch.distributed as dist
  from myriad_b200 import problems as PR
  from myriad_b200.config import Config, HParams, OptimizerType, QuadratureRule
  from myriad_b200.systems import SystemType
  from myriad_b200.trajectory_optimizers import get_optimizer
  from myriad_b200.utils import _rollout_engine

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py --impl b200 needs a CUDA device: myriad_b200 has no CPU fallback")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)

  hp = HParams(system=SystemType[SYSTEM], optimizer=OptimizerType.COLLOCATION, quadrature_rule=QuadratureRule[args.quadrature],
               intervals=INTERVALS, max_iter=1000, batch=args.batch)
  cfg = Config(verbose=False, plot=False)
  system = hp.system()
  opt = get_optimizer(hp, cfg, system)
  tr, eng = opt.transcription, opt.engine
  roll = _rollout_engine(hp, system)
  B = args.batch
  sz = eng.sizes

  # synthetic workload: global instance i = row i of one seeded draw; this rank owns rows [rank*B, (rank+1)*B)
  x0_all = PR.sample_x0(system, B * world, seed=hp.seed, spread=hp.start_spread)
  x0_host = x0_all[rank * B:(rank + 1) * B].contiguous().pin_memory()
  x0_dev = x0_host.to(dev)
  z0, lb, ub = PR.build_batch(tr, x0_dev)
  out = eng.ipm_solve(z0, lb, ub, max_iter=hp.max_iter)
  from myriad_b200.distributed import gather_solutions, pack_solution
  packed_w = sz.nvars + sz.ncon + 4
  gathered = torch.empty(B * world, packed_w, dtype=torch.float64, device=dev) if world > 1 else None
  flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # per-phase events of the last timed step (per-rank breakdown of the N>1 line)
  ph = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

  # The step's only collective -- the all_gather of the packed solutions -- runs on a side stream so that the gather of
  # step k overlaps the solve of step k + 1 (instances are independent: nothing in step k + 1 depends on it).  Every
  # timed step waits for the PREVIOUS step's gather before it ends, and the last timed step also for its own, so all K
  # gathers complete inside the timed region.
  side = torch.cuda.Stream(device=dev) if world > 1 else None
  gathered2 = [gathered, torch.empty_like(gathered)] if world > 1 else None
  pipe = {"k": 0, "pending": None}

  def device_step(last: bool = True):
    """inputs resident in HBM: solve + verification rollout + pack (+ gather, pipelined over steps when world > 1)"""
    ph[0].record()
    eng.ipm_solve(z0, lb, ub, max_iter=hp.max_iter, out=out)
    ph[1].record()
    _, u = tr.unravel(out["z"])
    _, cost = roll.rollout_cost(u.contiguous(), x0_dev, want_states=False)
    packed = pack_solution(out["z"], out["lam"], out["obj"], cost, out["status"], out["iters"])
    ph[2].record()
    if world == 1:
      ph[3].record()
      return packed
    main = torch.cuda.current_stream()
    if pipe["pending"] is not None:
      main.wait_event(pipe["pending"])      # gather of the previous step (ran under this step's solve)
    ready = torch.cuda.Event()
    ready.record(main)
    buf = gathered2[pipe["k"] & 1]
    pipe["k"] += 1
    with torch.cuda.stream(side):
      side.wait_event(ready)
      packed.record_stream(side)
      gather_solutions(packed, out=buf)
      done = torch.cuda.Event()
      done.record(side)
    pipe["pending"] = done
    if last:
      main.wait_event(done)
      pipe["pending"] = None
    ph[3].record()
    return buf

  nx = tr.nx_nodes * tr.n
  host_pack = torch.empty(B, packed_w, dtype=torch.float64).pin_memory()   # x, u (= xs_and_us), lambda, cost, rollout cost, status, iters

  def e2e_step():
    """public API from pinned host buffers: H2D(x0) -> build problem -> solve -> rollout -> D2H of the WHOLE result
    dictionary of the reference's solve() (x, u = xs_and_us; lambda; cost) + re-integrated cost, status, iterations"""
    x0 = x0_host.to(dev, non_blocking=True)
    sol = opt.solve_batch(x0)
    _, cost = roll.rollout_cost(sol["u"].contiguous(), x0, want_states=False)
    pk = pack_solution(sol["xs_and_us"], sol["lambda"], sol["cost"], cost, sol["status"], sol["iters"])
    host_pack.copy_(pk, non_blocking=True)
    gather_solutions(pk, out=gathered)
    torch.cuda.current_stream().synchronize()
    return x0.numel() * 8, host_pack.numel() * 8

  def timed(fn, steps, warmup, pipelined=False):
    for _ in range(warmup):
      fn()
    barrier()
    tot = 0.0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t_wall = time.time()
    for k, (a, b) in enumerate(ev):
      flush.fill_(1.0)  # L2 flush between timed iterations (outside the event pair)
      a.record()
      if pipelined:
        fn(last=(k == steps - 1))
      else:
        fn()
      b.record()
    barrier()
    wall = time.time() - t_wall
    tot = sum(a.elapsed_time(b) for a, b in ev)  # ms on the launching stream
    t = torch.tensor([tot], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), wall

  warm = max(args.warmup, 3)
  with ClockSampler(local) as clk:  # clocks are sampled during both timed regions (device-resident and end-to-end)
    ms_dev, wall_dev = timed(device_step, args.steps, warm, pipelined=True)
    ms_e2e, _ = timed(e2e_step, args.steps, warm)
  h2d, d2h = e2e_step()

  status = out["status"]
  n_ok = int((status == 0).sum())
  if world > 1:
    t = torch.tensor([n_ok], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    n_ok_all = int(t[0])
  else:
    n_ok_all = n_ok
  iters = out["iters"].double()

  # per-rank breakdown of the last timed device step (limiter of the multi-GPU scaling: VERDICT r1 weak #11)
  device_step(); torch.cuda.synchronize()
  mine = torch.tensor([ph[0].elapsed_time(ph[1]), ph[1].elapsed_time(ph[2]), ph[2].elapsed_time(ph[3]), float(out["iters"].max()),
                       float(out["iters"].double().sum())], dtype=torch.float64, device=dev)
  if world > 1:
    allr = torch.empty(world, 5, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allr, mine)
  else:
    allr = mine[None]
  per_rank = [{"rank": r, "solve_ms": float(v[0]), "rollout_pack_ms": float(v[1]), "gather_ms": float(v[2]), "max_iters": int(v[3]),
               "sum_iters": int(v[4])} for r, v in enumerate(allr.cpu())]

  # ---- roofline of K1 (rollout + defect + block Jacobian kernel), timed live over a working set larger than L2
  roof = roof_ipm = None
  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
      pass
    peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    Bk = max(B, 8192)  # working set must exceed the 126 MB L2: 43 KB/instance -> >= 355 MB at 8192 (no flush needed)
    zk = out["z"][:1].expand(Bk, -1).contiguous() + 0.01 * torch.randn(Bk, sz.nvars, dtype=torch.float64, device=dev)
    r = eng.eval(zk)
    for _ in range(3):
      eng.eval(zk, out=r)
    torch.cuda.synchronize()
    nrep = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(nrep):
      eng.eval(zk, out=r)
    b.record()
    torch.cuda.synchronize()
    t_k1 = a.elapsed_time(b) / nrep
    alg_bytes = Bk * 8 * (sz.nvars + sz.ncon + sz.jac_block_doubles + sz.nvars + 1)  # read z; write c, Jblk, grad, f
    ach = alg_bytes / (t_k1 * 1e-3) / 1e9
    roof = {"kernel": "eval_kernel (K1: rollout + defects + block Jacobian)", "bound": "hbm", "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak, "peak_source": which, "traffic": ncu_traffic_bytes(K1_PROFILE, "eval_kernel"), "batch": Bk,
            "bytes_per_instance": alg_bytes // Bk, "us_per_launch": t_k1 * 1e3,
            "l2": f"{nrep} back-to-back launches over a {alg_bytes / 1e6:.0f} MB working set (> 126 MB L2), no flush",
            "traffic_source": K1_PROFILE + " (dram read+write per launch at B=8192, parsed at run time; part of the written "
                              "lines is still dirty in L2 when the launch ends)",
            "note": "K1 is launched stand-alone here (myr_eval); inside the timed step the same node evaluation runs fused "
                    "inside ipm_kernel, the step's dominant kernel ('roofline')"}

    # ---- fp64 peak (DFMA loop) and the interior-point kernel's algorithmic FLOP/s
    import ctypes as C
    from myriad_b200 import _lib as ML
    scratch = torch.empty(148 * 8 * 1024, dtype=torch.float64, device=dev)
    it_d = 20000
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ML.check(ML.lib().myr_bench_dfma(148 * 8, 1000, C.c_void_p(scratch.data_ptr()), st))
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(); ML.check(ML.lib().myr_bench_dfma(148 * 8, it_d, C.c_void_p(scratch.data_ptr()), st)); b.record()
      torch.cuda.synchronize()
      best = min(best, a.elapsed_time(b))
    fp64_peak = 2.0 * 8 * it_d * 1024 * 148 * 8 / (best * 1e-3) / 1e12
    fl = ipm_flops_per_iteration(sz)
    tot_iters = float(out["iters"].double().sum())
    # time of the ipm kernel alone (no rollout / pack): CUDA events around one more solve
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.ipm_solve(z0, lb, ub, max_iter=hp.max_iter, out=out); b.record()
    torch.cuda.synchronize()
    t_ipm = a.elapsed_time(b)
    ach_f = fl["total"] * tot_iters / (t_ipm * 1e-3) / 1e12
    roof_ipm = {"kernel": "ipm_kernel (K3 loop: K1 node evaluation + K2 block-CR KKT solve + line search): the dominant kernel, ~98% of the step",
                "bound": "fp64", "achieved": ach_f,
                "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_f / fp64_peak,
                "peak_source": "measured live: DFMA loop, 8 independent chains/thread, 148x8 CTAs x 1024 threads (MEASURED_PEAKS.json "
                               "has no fp64 figure); the kernel is fp64-latency bound, neither 'hbm' nor 'tensor'",
                "traffic": ncu_traffic_bytes(IPM_PROFILE, "ipm_kernel"),
                "traffic_source": IPM_PROFILE + " (dram read+write of one launch at B=1024, parsed at run time)",
                "flops_per_iteration": fl, "iterations_total": tot_iters, "ms_per_launch": t_ipm,
                "note": "algorithmic flops (one KKT solve per iteration; inertia-correction retries and line-search "
                        "re-evaluations beyond the first are not counted)"}

  cpu = twin = parity = None
  if rank == 0 and not args.no_cpu_baseline:
    cores = os.cpu_count() or 1
    inst = args.cpu_instances or cores
    gpu_obj = out["obj"].cpu().numpy()
    gpu_ok = (out["status"] == 0).cpu().numpy()
    try:
      r = run_cpu_oracle(args.quadrature, inst, cores)
      cpu = {"value": r["solves_per_s"], "unit": "solves/s", "cores": cores, "kind": "port",
             "sample": f"{inst} instances of the same workload, one SciPy-SLSQP solve per core in parallel "
                       f"({r['seconds']:.1f} s wall; oracle restatement of the reference transcription)",
             "success": int(sum(r["success"])), "median_cost": sorted(r["costs"])[len(r["costs"]) // 2]}
      # parity on the sampled rows: same instances (row i of the seeded draw) solved by the CUDA path and by SLSQP
      both = [i for i in range(inst) if r["success"][i] and gpu_ok[i]]
      rel = [abs(gpu_obj[i] - r["costs"][i]) / max(1.0, abs(r["costs"][i])) for i in both]
      parity = {"rows": inst, "both_solved": len(both), "same_basin(|dobj|<=5e-5 rel)": int(sum(1 for v in rel if v <= 5e-5)),
                "max_rel_obj_diff": max(rel) if rel else None,
                "gpu_not_worse": int(sum(1 for i in both if gpu_obj[i] <= r["costs"][i] + 1e-6 * max(1.0, abs(r["costs"][i])))),
                "note": "SLSQP at SciPy's default ftol=1e-6 (the reference's setting); the IPM is converged to 1e-8"}
    except Exception as e:  # pragma: no cover
      cpu = {"value": None, "unit": "solves/s", "cores": cores, "kind": "port", "sample": f"failed: {e}"}
    try:
      twin = run_cpu_twin(eng, tr, z0, lb, ub, rows=min(B, 64 * cores))
      tw_obj = twin.pop("obj")
      k = len(tw_obj)
      twin["max_rel_obj_diff_vs_gpu"] = float(max(abs(tw_obj[i] - gpu_obj[i]) / max(1.0, abs(gpu_obj[i])) for i in range(k) if gpu_ok[i]))
    except Exception as e:  # pragma: no cover
      twin = {"value": None, "unit": "solves/s", "cores": cores, "kind": "port", "sample": f"failed: {e}"}

  if rank == 0:
    total = B * world
    line = {
      "metric": METRIC, "value": total * args.steps / (ms_dev * 1e-3), "unit": "solves/s", "n_gpus": world, "steps": args.steps,
      "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f64", "data": "synthetic",
      "config": {"workload": workload_name(args.quadrature), "batch_per_gpu": B, "max_iter": hp.max_iter, "tol": 1e-8, "nvars": sz.nvars, "ncon": sz.ncon,
                 "l2": "256 MB buffer written between timed steps (outside the timed events)",
                 "step": "myr_ipm_solve + myr_rollout_cost + pack" + (" + NCCL all_gather (side stream: the gather of step k "
                         "overlaps the solve of step k+1; all K gathers complete inside the timed region)" if world > 1 else ""),
                 "e2e_returns": "x, u (= xs_and_us), lambda, cost, re-integrated cost, status, iterations per instance"},
      "solved": n_ok_all, "instances": total, "success_rate": n_ok_all / total,
      "iters": {"min": int(iters.min()), "median": float(iters.median()), "max": int(iters.max())},
      "wall_s_timed_region": wall_dev,
      "e2e": {"value": total * args.steps / (ms_e2e * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
              "ms_per_step": ms_e2e / args.steps},
      "gpu_launches": 2 * args.steps,
      "clocks": clk.summary(),
      "roofline": roof_ipm,
      "roofline_k1": roof,
      "cpu_baseline": cpu,
      "cpu_baseline_twin": twin,
      "parity_sample": parity,
      "per_rank": per_rank,
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  a = parse()
  if a.impl == "reference":
    reference_arm(a)
  else:
    b200_arm(a)
