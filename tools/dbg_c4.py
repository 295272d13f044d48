import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("PIRIP_LDPC_STANDIN_OK", "1")
import torch
import bench_secondary as bs
import pirip_b200 as pb
dev = torch.device("cuda", 0)
ncap = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
iq, sent = bs.gen_coded_captures(torch, pb, ncap, 0, dev)
plan = pb.FskPlan(**bs.C4, soft=True, device=0)
nsamp = bs.C4_NSAMP
mf = plan.max_frames(nsamp)
llr = torch.zeros((ncap, mf * plan.nbits), dtype=torch.float32, device=dev)
nf = torch.zeros(ncap, dtype=torch.int32, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
fn = lambda: plan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), None, nf.data_ptr(), mf, d_llr=llr.data_ptr(), stream=sp)
r0 = plan.replays()
ms = bs.timed(torch, fn, 5)
r1 = plan.replays()
print("c4 captures", ncap, "ms %.3f" % ms, "flagged", plan.flagged(0)[0], "replays per call", (r1[0] - r0[0]) / 8.0, (r1[1] - r0[1]) / 8.0, flush=True)
