#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction stall samples with nvdisasm line info (-lineinfo build) and print
the hottest source lines.
Usage: tools/ncu_lines.py <report.ncu-rep> <object-or-so> <mangled-kernel-substring> <ncu -k regex> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, ksub, kre = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
linemap = {}
for cb in cubins:
  txt = subprocess.run(["nvdisasm", "-g", "-c", cb], stdout=subprocess.PIPE, text=True).stdout
  infn = False
  cur = None
  for ln in txt.splitlines():
    if ln.startswith("//---") and ".text." in ln:
      infn = ksub in ln
      continue
    if not infn:
      continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
      cur = (os.path.basename(m.group(1)), int(m.group(2)))
      continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
      linemap[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = collections.Counter(); inst = collections.Counter(); reasons = collections.defaultdict(collections.Counter)
RS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_lg", "stall_mio", "stall_math", "stall_no_inst",
      "stall_branch_resolving", "stall_dispatch", "stall_not_selected", "stall_selected", "stall_sleep", "stall_misc"]
totals = collections.Counter()
base = None
tot = 0
for r in rows:
  if r and r[0] == "Address":
    hdr = r; base = None; continue
  if hdr and len(r) == len(hdr):
    d = dict(zip(hdr, r))
    a = int(d["Address"], 16)
    if base is None:
      base = a
    s = int(d.get("# Samples") or 0)
    key = linemap.get(a - base, ("?", 0))
    agg[key] += s; tot += s
    inst[key] += int(d.get("Instructions Executed") or 0)
    for rname in RS:
      v = int(d.get(rname) or 0)
      reasons[key][rname] += v; totals[rname] += v
print(f"total samples {tot}")
print("by reason: " + ", ".join(f"{k[6:]} {100*v/max(tot,1):.1f}%" for k, v in totals.most_common(8)))
srcs = {}
for (f, l), s in agg.most_common(top):
  if f not in srcs:
    for root in ("myriad_b200/csrc", "."):
      pth = os.path.join(root, f)
      if os.path.exists(pth):
        srcs[f] = open(pth).read().splitlines(); break
    else:
      srcs[f] = []
  text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
  rs = ", ".join(f"{k[6:]} {100*v/max(s,1):.0f}%" for k, v in reasons[(f, l)].most_common(3))
  print(f"{100*s/max(tot,1):5.1f}% {inst[(f,l)]:>10d} inst  {f}:{l}  [{rs}]  {text}")
