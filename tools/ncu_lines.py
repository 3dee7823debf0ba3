#!/usr/bin/env python
"""Per-source-line stall-sample shares of an .ncu-rep: python tools/ncu_lines.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = [r for r in rows if r and r[0] == 'Line No'][0]; n = len(hdr)
cur = None; out = []; tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or not r[0].isdigit(): continue
    extra = len(r) - n
    src = ','.join(r[1:2 + extra]); vals = r[2 + extra:]
    try: s = int(vals[2]); ie = int(vals[5])
    except Exception: continue
    out.append((s, ie, cur, int(r[0]), src.strip())); tot += s
out.sort(reverse=True)
print("total samples", tot)
for s, ie, f, ln, src in out[:top]:
    print(f"{100*s/tot:5.1f}% inst={ie:>11} {f}:{ln}  {src[:105]}")

# optional grouping: python tools/ncu_lines.py rep top groups.txt   with lines "name file lo hi"
if len(sys.argv) > 3:
    groups = [l.split() for l in open(sys.argv[3]) if l.strip() and not l.startswith('#')]
    acc = {}
    for s, ie, f, ln, src in out:
        name = 'other'
        for g in groups:
            if f == g[1] and int(g[2]) <= ln <= int(g[3]): name = g[0]; break
        acc[name] = acc.get(name, 0) + s
    print("---- groups")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"{100*v/tot:6.1f}%  {k}")
