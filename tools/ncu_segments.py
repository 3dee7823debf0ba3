#!/usr/bin/env python
"""Segment a kernel's SASS by execution count and report stall-sample share and cycles per execution:
python tools/ncu_segments.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}; body=rows[2:]
S=[int(r[ix["# Samples"]]) for r in body]; IE=[int(r[ix["Instructions Executed"]]) for r in body]
sel=[int(r[ix["stall_selected"]]) for r in body]
tot=sum(S); unit=sum(sel)/sum(IE)
print("total samples",tot,"total warp-instr",sum(IE))
prev=None; start=0; segs=[]
for n,ie in enumerate(IE):
    if prev is None or abs(ie-prev)>0.02*max(prev,1):
        if prev is not None: segs.append((start,n,prev))
        start=n
    prev=ie
segs.append((start,len(IE),prev))
big=sorted(((sum(S[a:b]),a,b,ie) for a,b,ie in segs),reverse=True)
for s,a,b,ie in big[:top]:
    stl={}
    for r in body[a:b]:
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stl[h[6:]]=stl.get(h[6:],0)+int(r[ix[h]])
    tops=" ".join(f"{k}:{100*v/max(s,1):.0f}%" for k,v in sorted(stl.items(),key=lambda kv:-kv[1])[:4])
    print(f"rows {a:5d}-{b:5d} n={b-a:4d} execs={ie:>10} instr={100*(b-a)*ie/sum(IE):5.1f}% samples={100*s/tot:5.1f}% cyc/exec={s/max(ie,1)/unit:8.0f} {body[a][ix['Source']].strip()[:28]:28s} {tops}")
