#!/usr/bin/env python3
"""tools/sass_spills.py OBJ KERNEL_SUBSTR -- local-memory loads / stores (STL / LDL) of one kernel per source line (nvdisasm -g)."""
import collections, os, re, subprocess, sys, tempfile
obj, ksub = sys.argv[1], sys.argv[2]
td = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.split("\n")
start = end = None
for i, l in enumerate(txt):
    if l.startswith("\t.section\t.text.") and ksub in l and start is None: start = i
    elif l.startswith("\t.section") and start is not None and i > start: end = i; break
cur = None; cnt = collections.Counter()
for l in txt[start:end]:
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        k = re.search(r"\b(STL|LDL)\b", m.group(2))
        if k: cnt[(cur, k.group(1))] += 1
src = {}
for (f, ln), kind in sorted(cnt, key=lambda k: (k[0][0], k[0][1])):
    path = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
    if f not in src and os.path.exists(path): src[f] = open(path).read().split("\n")
    text = src[f][ln - 1].strip()[:100] if f in src else ""
    print(f"{f}:{ln} {kind} x{cnt[((f, ln), kind)]}  {text}")
