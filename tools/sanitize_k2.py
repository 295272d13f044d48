"""Small batch through the batch kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pirip_b200 as pb
import synth
ncap = int(sys.argv[1]) if len(sys.argv) > 1 else 300
caps = [synth.capture_u8(500 + i, 700, 2, 240000, 10000, 10000, 10000, ebno_db=[None, 9.0, 6.0][i % 3])[0] for i in range(6)]
n = min(len(c) for c in caps)
iq = np.stack([np.roll(caps[i % 6][:n], -2 * (i % 97)) for i in range(ncap)])
plan = pb.FskPlan(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000), device=0)
out = plan.demod_batch(iq, want_stats=True)
print("frames", out["nframes"][:8], "launches", pb.kernel_launches())
plan.close()
