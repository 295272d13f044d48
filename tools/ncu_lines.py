"""Map an ncu SASS source page to CUDA source lines using nvdisasm line info.

usage: ncu_lines.py <report.ncu-rep> <cubin> [top_n]
Prints instruction-executed and stall-sample shares per source line.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
# split per function (.text.<name> sections); pick the one whose size matches the SASS page
funcs = {}
curf = None
cur = None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        curf = m.group(1)
        funcs[curf] = []
        cur = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if curf and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l) and not any(x in l for x in (".byte", ".dword", ".short", ".word")):
        funcs[curf].append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ie, ss, sc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
inst = defaultdict(int)
smp = defaultdict(int)
body = rows[hi + 1:]
fname = min(funcs, key=lambda f: abs(len(funcs[f]) - len(body)))
line_of = funcs[fname]
print("function", fname[:80], len(line_of), "instructions; SASS rows", len(body))
for k, r in enumerate(body):
    key = line_of[k] if k < len(line_of) else None
    inst[key] += int(r[ie])
    smp[key] += int(r[ss])
ti, ts = sum(inst.values()), sum(smp.values())
print("total warp-instructions %d, stall samples %d" % (ti, ts))
src_cache = {}
for key in sorted(inst, key=lambda k: -(inst[k] if len(sys.argv) > 4 else smp[k]))[:top]:
    text = ""
    if key:
        f, n = key
        if f not in src_cache:
            try:
                src_cache[f] = open("pirip_b200/csrc/" + f).read().splitlines()
            except OSError:
                src_cache[f] = []
        if 0 < n <= len(src_cache[f]):
            text = src_cache[f][n - 1].strip()[:90]
    print("%5.1f%% inst %5.1f%% samples  %s  %s" % (100 * inst[key] / ti, 100 * smp[key] / ts, key, text))
