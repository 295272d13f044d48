"""Per source line: instructions executed, stall samples and the dominant stall reasons, from an ncu source-page CSV.

usage: ncu_stalls.py <src.csv> <cubin> [top_n] [line_lo line_hi]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lo = int(sys.argv[4]) if len(sys.argv) > 5 else None
hi_ = int(sys.argv[5]) if len(sys.argv) > 5 else None
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
funcs, curf, cur = {}, None, None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        curf = m.group(1); funcs[curf] = []; cur = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if curf and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l) and not any(x in l for x in (".byte", ".dword", ".short", ".word")):
        funcs[curf].append(cur)
rows = list(csv.reader(open(src)))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr, body = rows[h], rows[h + 1:]
fname = min(funcs, key=lambda f: abs(len(funcs[f]) - len(body)))
lines = funcs[fname]
print("function", fname[:60], len(lines), "instr; rows", len(body))
ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
inst, smp, why = defaultdict(int), defaultdict(int), defaultdict(lambda: defaultdict(int))
for k, r in enumerate(body):
    ln = lines[k] if k < len(lines) else None
    inst[ln] += int(r[ie] or 0); smp[ln] += int(r[ns] or 0)
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: why[ln][hdr[c][6:]] += v
ti, ts = sum(inst.values()), sum(smp.values())
print("total warp-instr", ti, "samples", ts)
keys = [k for k in inst if k and (lo is None or (k[0].startswith("fsk_kernel2") and lo <= k[1] <= hi_))]
keys.sort(key=lambda k: -smp[k])
for k in keys[:top]:
    w = sorted(why[k].items(), key=lambda kv: -kv[1])[:4]
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * inst[k] / ti, 100.0 * smp[k] / ts, k[0], k[1], " ".join("%s=%d" % kv for kv in w)))
