#!/usr/bin/env python
"""Static SASS instruction count per source line for one kernel (nvdisasm -g line info).
Usage: tools/code_size_by_line.py <obj> <mangled-kernel-substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, ksub = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cnt = collections.Counter(); tot = 0
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
      cnt[cur] += 1; tot += 1
print("total instructions", tot, "=", tot * 16 // 1024, "KB")
byfile = collections.Counter()
for (f, l), c in cnt.items(): byfile[f] += c
for f, c in byfile.most_common(): print(f"  {f}: {c} ({100*c/tot:.1f}%)")
srcs = {}
for (f, l), c in cnt.most_common(top):
  if f not in srcs:
    pth = os.path.join("myriad_b200/csrc", f)
    srcs[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
  text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
  print(f"{c:6d} {100*c/tot:5.1f}%  {f}:{l}  {text}")
