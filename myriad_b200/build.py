"""In-tree build of libmyriad_b200.so for sm_100a with nvcc (no JIT cache, no pip install).

One object per system (csrc/sys_unit.cu with -DMYR_SYS_CLASS=...) compiled in parallel, plus the thin
dispatcher csrc/api.cu; objects live in build/ and are reused when older than no source.

    python -m myriad_b200.build [--force] [--systems CARTPOLE,VANDERPOL] [--jobs 8]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(ROOT, "build")
LIB = os.path.join(HERE, "libmyriad_b200.so")

# true systems that also get a NodeSystem wrapper (neural-ODE dynamics on the tensor-core path, csrc/node_mlp.cuh).
# Each wrapper is one more translation unit (~3 CPU-minutes); BASELINE config C5 needs CARTPOLE.  Others:
#   MYR_NODE_SYSTEMS=CARTPOLE,VANDERPOL,CANCERTREATMENT python -m myriad_b200.build
NODE_SYSTEMS = [s for s in os.environ.get("MYR_NODE_SYSTEMS", "CARTPOLE").split(",") if s]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fopenmp"]


def _nvcc() -> str:
  for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError("nvcc not found: the CUDA toolkit is required to build myriad_b200")


def generated_systems():
  """Struct names in csrc/systems_gen.cuh -> {NAME: SysName}"""
  txt = open(os.path.join(CSRC, "systems_gen.cuh")).read()
  out = {}
  for cls, name in re.findall(r'struct (Sys\w+) \{[^}]*?static constexpr const char\* name = "(\w+)";', txt, flags=re.S):
    out[name] = cls
  return out


def _newest_source(exclude=()) -> float:
  t = 0.0
  for d in (CSRC, os.path.join(ROOT, "include")):
    for f in os.listdir(d):
      if f not in exclude:
        t = max(t, os.path.getmtime(os.path.join(d, f)))
  return t


def _run(cmd):
  r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  if r.returncode != 0:
    raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + r.stdout)
  return r.stdout


def build(force: bool = False, systems=None, jobs: int | None = None, verbose: bool = True, extra_flags=()) -> str:
  nvcc = _nvcc()
  os.makedirs(BUILD, exist_ok=True)
  gen = generated_systems()
  names = list(gen) if not systems else [s.upper() for s in systems]
  for s in names:
    if s not in gen:
      raise KeyError(f"system {s} has no generated device code (tools/gen_systems.py)")
  src_t = _newest_source()
  sys_src_t = _newest_source(exclude=("api.cu", "fbsm.cu", "fbsm.cuh"))  # the per-system units include neither
  tag = "_".join(sorted(names)) + "|node:" + "_".join(sorted(s for s in NODE_SYSTEMS if s in names))
  stamp = os.path.join(BUILD, "systems.txt")
  prev = open(stamp).read() if os.path.exists(stamp) else ""
  jobs = jobs or min(8, os.cpu_count() or 1)
  flags = NVCC_FLAGS + list(extra_flags)
  xmacro = "-DMYR_BUILD_SYSTEMS(X)=" + " ".join(f"X({gen[s]})" for s in names)
  tasks = []
  objs = []
  for s in names:
    obj = os.path.join(BUILD, f"sys_{s}.o")
    objs.append(obj)
    if force or not os.path.exists(obj) or os.path.getmtime(obj) < sys_src_t:
      tasks.append((s, [nvcc, *flags, f"-DMYR_SYS_CLASS={gen[s]}", "-c", os.path.join(CSRC, "sys_unit.cu"), "-o", obj]))
  node_names = [s for s in NODE_SYSTEMS if s in names]
  for s in node_names:
    obj = os.path.join(BUILD, f"node_{s}.o")
    objs.append(obj)
    if force or not os.path.exists(obj) or os.path.getmtime(obj) < sys_src_t:
      tasks.append(("node_" + s, [nvcc, *flags, f"-DMYR_SYS_CLASS={gen[s]}", "-DMYR_NODE=1", "-c", os.path.join(CSRC, "sys_unit.cu"), "-o", obj]))
  xmacro_node = "-DMYR_BUILD_NODE_SYSTEMS(X)=" + " ".join(f"X({gen[s]})" for s in node_names)
  api_obj = os.path.join(BUILD, "api.o")
  objs.append(api_obj)
  if force or not os.path.exists(api_obj) or os.path.getmtime(api_obj) < src_t or prev != tag:
    tasks.append(("api", [nvcc, *flags, xmacro, xmacro_node, "-c", os.path.join(CSRC, "api.cu"), "-o", api_obj]))
  fbsm_obj = os.path.join(BUILD, "fbsm.o")  # forward-backward sweep: needs only the generated systems
  objs.append(fbsm_obj)
  fbsm_t = max(os.path.getmtime(os.path.join(CSRC, f)) for f in ("fbsm.cu", "fbsm.cuh", "systems_gen.cuh"))
  fbsm_t = max(fbsm_t, os.path.getmtime(os.path.join(ROOT, "include", "myriad_b200.h")))
  if force or not os.path.exists(fbsm_obj) or os.path.getmtime(fbsm_obj) < fbsm_t:
    tasks.append(("fbsm", [nvcc, *flags, "-c", os.path.join(CSRC, "fbsm.cu"), "-o", fbsm_obj]))
  if tasks:
    if verbose:
      print(f"[myriad_b200.build] compiling {len(tasks)} unit(s) with {jobs} job(s): {[t[0] for t in tasks]}", flush=True)
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
      for out in ex.map(lambda t: _run(t[1]), tasks):
        if verbose and out.strip():
          print(out)
  if tasks or not os.path.exists(LIB) or prev != tag:
    _run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fopenmp", "-lgomp"])
    open(stamp, "w").write(tag)
    if verbose:
      print(f"[myriad_b200.build] linked {LIB}", flush=True)
  return LIB


if __name__ == "__main__":
  ap = argparse.ArgumentParser()
  ap.add_argument("--force", action="store_true")
  ap.add_argument("--systems", default=None)
  ap.add_argument("--jobs", type=int, default=None)
  a = ap.parse_args()
  build(force=a.force, systems=a.systems.split(",") if a.systems else None, jobs=a.jobs)
  sys.exit(0)
