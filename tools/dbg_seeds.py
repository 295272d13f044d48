"""Which seeds of the bench workload leave captures to the exact kernel, and why."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
import pirip_b200 as pb
dev = torch.device("cuda", 0)
nsamp = 240000
plan = pb.FskPlan(bench.FS, bench.RS, bench.M, P=bench.P, Nsym=bench.NSYM, limits=bench.LIMITS, device=0)
mf = plan.max_frames(nsamp)
ncap = 1250
bits = torch.zeros((ncap, mf * plan.nbits), dtype=torch.uint8, device=dev)
nf = torch.zeros(ncap, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream()
for seed in [int(x) for x in sys.argv[1:]] or range(8):
    iq = bench.gen_captures(torch, ncap, nsamp, seed, dev)
    r0 = plan.replays()
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        plan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), bits.data_ptr(), nf.data_ptr(), mf, stream=ctypes.c_void_p(s.cuda_stream))
        e1.record(s)
        torch.cuda.synchronize()
    cnt, fl = plan.flagged(ncap)
    r1 = plan.replays()
    idx = np.nonzero(fl)[0]
    print("seed", seed, "ms %.3f" % e0.elapsed_time(e1), "flagged", cnt, "replays/call", (r1[0] - r0[0]) / 3, (r1[1] - r0[1]) / 3,
          [(int(i), hex(int(fl[i]) & 0xff), (int(fl[i]) >> 8) & 0xfff, i % 4) for i in idx[:6]], flush=True)
    del iq
