"""A/B timing of batch-kernel variants selected by environment variables, alternating in one process-per-run loop.
usage: python tools/ab_k2.py VAR=a,b,c [--captures N] [--reps R]"""
import os
import subprocess
import sys

var, vals = sys.argv[1].split("=")
vals = vals.split(",")
caps = [a for a in sys.argv[2:] if a.isdigit()] or ["1250"]
reps = 3
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for c in caps:
    res = {v: [] for v in vals}
    for r in range(reps):
        for v in vals:
            env = dict(os.environ)
            env[var] = v
            out = subprocess.run([sys.executable, os.path.join(root, "tools", "prof_fsk.py"), "--captures", c, "--seconds", "1.0", "--iters", "5"],
                                 env=env, capture_output=True, text=True).stdout
            res[v].append(float(out.split(":")[1].split("ms")[0]))
    for v in vals:
        print("captures %s %s=%s ms %s" % (c, var, v, " ".join("%.3f" % x for x in res[v])))
