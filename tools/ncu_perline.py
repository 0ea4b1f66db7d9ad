#!/usr/bin/env python
"""Instructions per source line of the profiled kernel, normalised by `units` (e.g. warp row-iterations).
usage: python tools/ncu_perline.py report.ncu-rep units [min]"""
import csv, subprocess, sys
rep=sys.argv[1]; units=float(sys.argv[2]); mn=float(sys.argv[3]) if len(sys.argv)>3 else 2.0
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None;hdr=None;lines=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Line No": hdr=r; ii=hdr.index("Instructions Executed"); si=hdr.index("# Samples"); continue
    if r[0] in ("Function Name","") or hdr is None: continue
    try: lines.append((cur,int(r[0]),r[1].strip(),int(r[ii]),int(r[si])))
    except: pass
tot=sum(x[3] for x in lines); ts=sum(x[4] for x in lines) or 1
print("total",tot, "per-unit", tot/units)
for f,ln,src,n,sm in sorted(lines,key=lambda x:(x[0],x[1])):
    v=n/units
    if v>=mn or sm/ts>0.015: print(f"{v:6.1f} {100*sm/ts:5.1f}%  {f}:{ln}: {src[:105]}")
