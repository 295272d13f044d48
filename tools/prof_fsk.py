"""Small driver for ncu: runs the FSK demod kernel a few times on a reduced batch."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import pirip_b200 as pb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--captures", type=int, default=592)
ap.add_argument("--seconds", type=float, default=0.25)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
nsamp = int(a.seconds * bench.FS)
plan = pb.FskPlan(bench.FS, bench.RS, bench.M, P=bench.P, Nsym=bench.NSYM, limits=bench.LIMITS, device=0)
mf = plan.max_frames(nsamp)
iq = bench.gen_captures(torch, a.captures, nsamp, 0, dev)
bits = torch.zeros((a.captures, mf * plan.nbits), dtype=torch.uint8, device=dev)
nf = torch.zeros(a.captures, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters):
    if i == a.iters - 1:
        e0.record(s)
    plan.demod_batch_dev(iq.data_ptr(), a.captures, nsamp, iq.stride(0), bits.data_ptr(), nf.data_ptr(), mf,
                         stream=ctypes.c_void_p(s.cuda_stream))
e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("captures %d x %d samples: %.3f ms, %.1f Msamples/s" % (a.captures, nsamp, ms, a.captures * nsamp / ms / 1e3))
import numpy as np
cnt, fl = plan.flagged(a.captures)
print("mode", plan.mode, "flagged", cnt, "reasons", {k: int(((fl & v) != 0).sum()) for k, v in (("sym", 1), ("tim", 2), ("ppm", 4), ("nan", 8))},
      "first-frame-0:", int((((fl >> 8) & 0xfff) == 1).sum()), "replays(ok,failed)", plan.replays(), plan.kernel())
