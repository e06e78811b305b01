#!/usr/bin/env python3
"""tools/sass_lines.py OBJ KERNEL_SUBSTR [FILE:LO-HI ...] -- static SASS instruction counts of one kernel per source file (nvdisasm -g line
info), the opcode histogram, and (optionally) the listing of the instructions attributed to a source-line range.  CPU only (no GPU)."""
import collections, os, re, subprocess, sys, tempfile


def main():
    obj, ksub = sys.argv[1], sys.argv[2]
    ranges = []
    for a in sys.argv[3:]:
        f, r = a.split(":"); lo, hi = r.split("-"); ranges.append((f, int(lo), int(hi)))
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.split("\n")
    start = end = None
    for i, l in enumerate(txt):
        if l.startswith("\t.section\t.text.") and ksub in l and start is None:
            start = i
        elif l.startswith("\t.section") and start is not None and i > start:
            end = i; break
    lines = txt[start:end]
    cur = None; byfile = collections.Counter(); ops = collections.Counter(); listing = []
    for l in lines:
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            byfile[cur[0] if cur else "?"] += 1
            op = re.sub(r"^@!?U?P[0-9T]+\s+", "", m.group(2)).split()[0].split(".")[0]
            ops[op] += 1
            for f, lo, hi in ranges:
                if cur and cur[0] == f and lo <= cur[1] <= hi:
                    listing.append(f"{cur[1]:4d} {m.group(1)} {m.group(2)}")
    print("total", sum(byfile.values()), dict(byfile))
    print("ops", ops.most_common(24))
    if ranges:
        print(len(listing), "instructions in ranges")
        print("\n".join(listing))


if __name__ == "__main__":
    main()
