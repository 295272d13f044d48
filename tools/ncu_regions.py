"""Instruction and stall-sample shares per SASS region of an ncu report (chunks of N instructions in program order, with the
source lines and opcodes that dominate each chunk): shows which PHASE of a long kernel the instructions belong to when the
per-line view is blurred by inlined helpers.

usage: ncu_regions.py <report.ncu-rep> <cubin> [chunk=250]
"""
import csv, re, subprocess, sys
from collections import Counter
rep, cubin = sys.argv[1], sys.argv[2]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 250
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
funcs = {}; curf=None; cur=None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m: curf=m.group(1); funcs[curf]=[]; cur=None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    if curf and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l) and not any(x in l for x in (".byte",".dword",".short",".word")):
        funcs[curf].append(cur)
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=next(i for i,r in enumerate(rows) if "Instructions Executed" in r)
hdr=rows[hi]; ie,ss,sc=hdr.index("Instructions Executed"),hdr.index("# Samples"),hdr.index("Source")
body=rows[hi+1:]
fname=min(funcs,key=lambda f:abs(len(funcs[f])-len(body)))
lo=funcs[fname]
ti=sum(int(r[ie]) for r in body); ts=sum(int(r[ss]) for r in body)
print("total",ti,ts)
for k0 in range(0,len(body),chunk):
    seg=body[k0:k0+chunk]
    i=sum(int(r[ie]) for r in seg); s=sum(int(r[ss]) for r in seg)
    c=Counter()
    for k,r in enumerate(seg):
        key=lo[k0+k] if k0+k<len(lo) else None
        if key and key[0]=='fsk_fast.cu': c[key[1]]+=int(r[ie])
    top=[str(l) for l,_ in c.most_common(6)]
    ops=Counter(r[sc].split()[0] if not r[sc].startswith('@') else r[sc].split()[1] for r in seg)
    print("%5d %5.1f%% inst %5.1f%% smp lines %s ops %s"%(k0,100*i/ti,100*s/ts,",".join(top)," ".join(f"{o}:{n}" for o,n in ops.most_common(5))))
