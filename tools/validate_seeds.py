"""Fast mode against the exact kernel (itself bit-identical to the oracle, tests/test_fsk_gpu*.py) on fresh batches of the bench
workload: hard bits and frame counts of every capture must be equal.  usage: validate_seeds.py [first_seed [n_seeds [captures]]]   (SEEDS=a,b,c: an explicit list)"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import pirip_b200 as pb

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
nseeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ncap = int(sys.argv[3]) if len(sys.argv) > 3 else 1250
seeds = [int(x) for x in os.environ["SEEDS"].split(",")] if os.environ.get("SEEDS") else list(range(first, first + nseeds))
dev = torch.device("cuda", 0)
nsamp = 240000
fast = pb.FskPlan(bench.FS, bench.RS, bench.M, P=bench.P, Nsym=bench.NSYM, limits=bench.LIMITS, device=0)
exact = pb.FskPlan(bench.FS, bench.RS, bench.M, P=bench.P, Nsym=bench.NSYM, limits=bench.LIMITS, device=0)
exact.set_mode("exact")
mf = fast.max_frames(nsamp)
s = torch.cuda.current_stream()
tot = dict(captures=0, captures_mismatching=0, bit_mismatches=0, frame_count_mismatches=0, replays=0, left_to_exact_kernel=0)
for seed in seeds:
    iq = bench.gen_captures(torch, ncap, nsamp, seed, dev)
    out = []
    for plan in (fast, exact):
        bits = torch.zeros((ncap, mf * plan.nbits), dtype=torch.uint8, device=dev)
        nf = torch.zeros(ncap, dtype=torch.int32, device=dev)
        r0 = plan.replays() if plan is fast else None
        plan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), bits.data_ptr(), nf.data_ptr(), mf, stream=ctypes.c_void_p(s.cuda_stream))
        torch.cuda.synchronize()
        if plan is fast:
            tot["replays"] += plan.replays()[0] - r0[0]
            tot["left_to_exact_kernel"] += int(plan.flagged(ncap)[0])
        out.append((bits, nf))
    (b0, n0), (b1, n1) = out
    bad_nf = int((n0 != n1).sum())
    diff = (b0 != b1)
    bad_caps = int((diff.any(dim=1) | (n0 != n1)).sum())
    tot["captures"] += ncap
    tot["captures_mismatching"] += bad_caps
    tot["bit_mismatches"] += int(diff.sum())
    tot["frame_count_mismatches"] += bad_nf
    where = [(int(c), int(diff[c].nonzero()[0, 0]) // fast.nbits, int(diff[c].sum())) for c in diff.any(dim=1).nonzero()[:, 0][:4]]
    print(json.dumps(dict(seed=seed, captures_mismatching=bad_caps, bit_mismatches=int(diff.sum()), frame_count_mismatches=bad_nf,
                          first=[dict(capture=c, frame=f, bits=b) for c, f, b in where])), flush=True)
    del iq
print(json.dumps(dict(total=tot, bits_per_capture=int(mf * fast.nbits))))
