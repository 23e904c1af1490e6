#!/usr/bin/env python
"""Static count of SASS instructions matching a regex per source line for one kernel (nvdisasm -g line info).
Usage: tools/sass_by_line.py <obj> <mangled-kernel-substring> <regex> [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, ksub, rx = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cnt = collections.Counter(); tot = 0; alli = 0
for f in os.listdir(tmp):
  if not f.endswith(".cubin"): continue
  txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
  infn = False; cur = None
  for ln in txt.splitlines():
    if ln.startswith("//---") and ".text." in ln:
      infn = ksub in ln; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
      cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln) and cur:
      alli += 1
      if re.search(rx, ln): cnt[cur] += 1; tot += 1
print("matching", tot, "of", alli)
srcs = {}
for (f, l), c in cnt.most_common(top):
  if f not in srcs:
    pth = os.path.join("myriad_b200/csrc", f)
    srcs[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
  text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
  print(f"{c:6d}  {f}:{l}  {text}")
