"""Fast mode against the exact kernel (itself bit-identical to the oracle, tests/test_fsk_gpu*.py) on fresh batches of the bench
workload: hard bits and frame counts of every capture must be equal.  usage: validate_seeds.py [first_seed [n_seeds [captures]]]   (SEEDS=a,b,c: an explicit list; CONFIG=c3: the 4FSK shape of config 3)"""
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
import numpy as np


def gen_4fsk(torch, ncap, nsamp, seed, dev):
    """BASELINE.json config 3 (4FSK Fs=240k Rs=10k P=8): the generator of tools/bench_secondary.py::c3_4fsk_llr with a seed"""
    Fs, Rs = 240000, 10000
    g = torch.Generator(device=dev)
    g.manual_seed(0x50495250 + 333 + 7919 * seed)
    frame = torch.tensor([int(c) for c in bench.TEST_FRAME], dtype=torch.int64, device=dev)
    syms = frame[0::2] * 2 + frame[1::2]
    Ts = Fs // Rs
    iq = torch.empty((ncap, 2 * nsamp), dtype=torch.uint8, device=dev)
    n = torch.arange(nsamp, dtype=torch.float64, device=dev)[None, :]
    ebnos = (None, 12.0, 9.0, 6.0)
    for c0 in range(0, ncap, 50):
        nc = min(50, ncap - c0)
        u = torch.rand((nc, 4), generator=g, device=dev, dtype=torch.float64)
        toff = u[:, 1:2] * Ts
        ts_eff = Ts * (1.0 + (u[:, 3:4] * 2 - 1) * 50e-6)
        k = torch.floor((n - toff) / ts_eff).clamp_(min=0).to(torch.int64)
        sym = syms[k % syms.numel()].to(torch.float64)
        f = (937.5 + (u[:, 2:3] * 2 - 1) * 0.02 * Rs + 7500.0 * sym) / Fs
        ph = u[:, 0:1] * 2 * np.pi + 2 * np.pi * torch.cumsum(f, dim=1)
        eb = torch.tensor([ebnos[(c0 + i) % 4] or 1e9 for i in range(nc)], dtype=torch.float64, device=dev)[:, None]
        sigma = torch.sqrt(Ts / (2.0 * 2.0 * 10 ** (eb / 10.0)))
        amp = torch.clamp(110.0 / (1.0 + 3.0 * sigma), max=100.0)
        re = torch.cos(ph) + sigma * torch.randn((nc, nsamp), generator=g, device=dev, dtype=torch.float64)
        im = torch.sin(ph) + sigma * torch.randn((nc, nsamp), generator=g, device=dev, dtype=torch.float64)
        iq[c0:c0 + nc, 0::2] = torch.clamp(torch.round(127.5 + amp * re), 0, 255).to(torch.uint8)
        iq[c0:c0 + nc, 1::2] = torch.clamp(torch.round(127.5 + amp * im), 0, 255).to(torch.uint8)
    return iq


if os.environ.get("CONFIG") == "c3":  # CONFIG=c3: 4FSK, P = 8
    pk = dict(Fs=240000, Rs=10000, M=4, P=8, Nsym=50, limits=(500, 25000))
    gen = gen_4fsk
else:  # the bench workload; CONFIG=p8: the same captures demodulated with P = 8 (the shape of config 4's demodulator)
    pk = dict(Fs=bench.FS, Rs=bench.RS, M=bench.M, P=8 if os.environ.get("CONFIG") == "p8" else bench.P, Nsym=bench.NSYM, limits=bench.LIMITS)
    gen = bench.gen_captures
fast = pb.FskPlan(pk["Fs"], pk["Rs"], pk["M"], P=pk["P"], Nsym=pk["Nsym"], limits=pk["limits"], device=0)
exact = pb.FskPlan(pk["Fs"], pk["Rs"], pk["M"], P=pk["P"], Nsym=pk["Nsym"], limits=pk["limits"], device=0)
exact.set_mode("exact")
mf = fast.max_frames(nsamp)
s = torch.cuda.current_stream()
tot = dict(captures=0, captures_mismatching=0, bit_mismatches=0, frame_count_mismatches=0, replays=0, left_to_exact_kernel=0)
for seed in seeds:
    iq = gen(torch, ncap, nsamp, seed, dev)
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
