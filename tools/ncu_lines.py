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
line_of = []  # per instruction, in order
cur = None
for l in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l) and ".byte" not in l and ".dword" not in l and ".short" not in l and ".word" not in l:
        line_of.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ie, ss, sc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
inst = defaultdict(int)
smp = defaultdict(int)
body = rows[hi + 1:]
if len(body) != len(line_of):
    print("warning: %d SASS rows vs %d disassembled instructions" % (len(body), len(line_of)))
for k, r in enumerate(body):
    key = line_of[k] if k < len(line_of) else None
    inst[key] += int(r[ie])
    smp[key] += int(r[ss])
ti, ts = sum(inst.values()), sum(smp.values())
print("total warp-instructions %d, stall samples %d" % (ti, ts))
src_cache = {}
for key in sorted(inst, key=lambda k: -smp[k])[:top]:
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
