"""Who is right when the fast and the exact kernel disagree: the oracle (tests/orc.py) on single captures of the bench workload.
usage: arbitrate.py seed:capture [seed:capture ...]   (test tooling: imports oracle/ through tests/orc.py)"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
import orc
import pirip_b200 as pb

dev = torch.device("cuda", 0)
nsamp, ncap = 240000, 1250
fast = pb.FskPlan(bench.FS, bench.RS, bench.M, P=bench.P, Nsym=bench.NSYM, limits=bench.LIMITS, device=0)
exact = pb.FskPlan(bench.FS, bench.RS, bench.M, P=bench.P, Nsym=bench.NSYM, limits=bench.LIMITS, device=0)
exact.set_mode("exact")
cfg = orc.demod_cfg(bench.FS, bench.RS, bench.M, P=bench.P, Nsym=bench.NSYM, limits=bench.LIMITS)
mf = fast.max_frames(nsamp)
s = torch.cuda.current_stream()
for arg in sys.argv[1:]:
    seed, cap = (int(x) for x in arg.split(":"))
    iq = bench.gen_captures(torch, ncap, nsamp, seed, dev)
    res, stats = {}, {}
    for name, plan in (("fast", fast), ("exact", exact)):
        bits = torch.zeros((ncap, mf * plan.nbits), dtype=torch.uint8, device=dev)
        nf = torch.zeros(ncap, dtype=torch.int32, device=dev)
        st = torch.zeros((ncap, mf, 20), dtype=torch.float32, device=dev)  # struct pirip_frame_stats: 18 floats + 2 int32
        r0 = plan.replays()
        plan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), bits.data_ptr(), nf.data_ptr(), mf, d_stats=st.data_ptr(),
                             stream=ctypes.c_void_p(s.cuda_stream))
        torch.cuda.synchronize()
        res[name] = (bits[cap].cpu().numpy(), int(nf[cap]))
        stats[name] = st[cap].cpu().numpy()
    ref = orc.fsk_demod_stream(cfg, iq[cap].cpu().numpy())
    out = dict(seed=seed, capture=cap, oracle_frames=int(ref["nframes"]))
    for name, (b, n) in res.items():
        m = min(n, int(ref["nframes"])) * fast.nbits
        d = np.nonzero(b[:m] != ref["bits"][:m])[0]
        out[name] = dict(frames=n, bit_mismatches_vs_oracle=int(d.size), first_frame=int(d[0]) // fast.nbits if d.size else None)
    print(json.dumps(out), flush=True)
    ff = out["fast"]["first_frame"]
    if ff is not None:
        for f in range(max(ff - 1, 0), min(ff + 2, mf)):
            row = {}
            for name in ("fast", "exact"):
                r = stats[name][f]
                row[name] = dict(norm_rx_timing=float(r[8]), rx_timing=float(r[17]), nin_next=int(r[18:19].view(np.int32)[0]),
                                 offset=int(r[19:20].view(np.int32)[0]), f_est=[float(r[0]), float(r[1])])
            b, n = res["fast"]
            d = np.nonzero(b[f * fast.nbits:(f + 1) * fast.nbits] != ref["bits"][f * fast.nbits:(f + 1) * fast.nbits])[0]
            print("  frame", f, json.dumps(row), "wrong bits", d.tolist(), flush=True)
    del iq
