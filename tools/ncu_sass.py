#!/usr/bin/env python
"""SASS-level stall samples of an .ncu-rep: python tools/ncu_sass.py rep.ncu-rep [lo hi]  (row range) | --find PATTERN"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
if len(sys.argv) > 2 and sys.argv[2] == "--find":
    for n, r in enumerate(body):
        if sys.argv[3] in r[ix["Source"]]:
            print(n, r[ix["Source"]].strip())
    sys.exit()
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(body)
print("total samples", tot, "rows", len(body))
for n in range(lo, min(hi, len(body))):
    r = body[n]
    s = int(r[ix["# Samples"]]); ie = int(r[ix["Instructions Executed"]])
    top = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    ts = " ".join(f"{nm}:{v}" for v, nm in top if v)
    print(f"{n:5d} {100*s/tot:5.2f}% {s:7d} ie={ie:>10} spi={s/max(ie,1)*1e3:7.3f} {r[ix['Source']].strip()[:70]:70s} {ts}")
