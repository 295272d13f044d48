"""Diagnosis: fast kernel with more captures than one round of resident groups."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import pirip_b200 as pb
cfgname, ncap, seconds = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
kw = {"c1": dict(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000)),
      "c4": dict(Fs=240000, Rs=10000, M=2, P=8, limits=(500, 25000), soft=True)}[cfgname]
dev = torch.device("cuda", 0)
nsamp = int(seconds * 240000)
plan = pb.FskPlan(**kw, device=0)
mf = plan.max_frames(nsamp)
iq = bench.gen_captures(torch, ncap, nsamp, 0, dev, ebno_list=(None, 12.0, 6.0, 3.0))
bits = torch.zeros((ncap, mf * plan.nbits), dtype=torch.uint8, device=dev)
llr = torch.zeros((ncap, mf * plan.nbits), dtype=torch.float32, device=dev)
nf = torch.zeros(ncap, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream()
for i in range(3):
    t0 = time.time()
    plan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), bits.data_ptr(), nf.data_ptr(), mf, d_llr=llr.data_ptr() if kw.get("soft") else None,
                         stream=ctypes.c_void_p(s.cuda_stream))
    torch.cuda.synchronize()
    print(cfgname, ncap, "call", i, "%.1f ms" % ((time.time() - t0) * 1e3), "flagged", plan.flagged(0)[0], "replays", plan.replays(), flush=True)
print(plan.kernel())
