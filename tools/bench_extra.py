#!/usr/bin/env python
"""Secondary measurements (not the headline bench): the fused csdr stage and the LDPC decoder.

  decim : config-2 front end, u8 IQ at 1.8 MS/s -> convert + 79-tap FIR + /45 + s16, streams
          resident in HBM; algorithmic bytes = 2 B per input sample + 4 B per output
  ldpc  : H_256_512_4-sized codewords at a mix of SNRs, LLRs resident in HBM;
          algorithmic bytes = 4*N in + N out per codeword (compute-bound, reported for completeness)
Prints one JSON object per workload.  CUDA events on torch's current stream, inputs >> L2.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import pirip_b200 as pb  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (json.load(open(p))["hbm_gbs"], "measured") if os.path.exists(p) else (6650.0, "fallback")


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_decim(steps, D=45, n=1800000):
    nstreams = 256  # 256 one-second streams at 1.8 MS/s (D = 45) or 2.4 MS/s (D = 30): 0.9 / 1.2 GB of u8
    dec = pb.DecimPlan(D)
    g = torch.Generator(device="cuda").manual_seed(1)
    iq = torch.randint(0, 256, (nstreams, 2 * n), dtype=torch.uint8, device="cuda", generator=g)
    nout = dec.out_count(n)
    out = torch.empty((nstreams, 2 * nout), dtype=torch.int16, device="cuda")
    s = torch.cuda.current_stream()
    fn = lambda: dec.run_dev(iq.data_ptr(), nstreams, n, iq.stride(0), None, out.data_ptr(), nout,
                             stream=ctypes.c_void_p(s.cuda_stream))
    ms = timed(fn, steps)
    # parity spot check against the oracle on one stream prefix
    import orc
    ref_f, ref_s = orc.csdr_decimate_stream(iq[3, :2 * 100000].cpu().numpy(), D)
    ok = bool(np.array_equal(out[3, :ref_s.size].cpu().numpy(), ref_s))
    algo = nstreams * (2.0 * n + 4.0 * nout)
    pk, src = peak()
    ach = algo / (ms * 1e-3) / 1e9
    return {"workload": "decim: %d x %.1f Msample u8 IQ streams, fir_decimate_cc %d (%d taps), s16 out" % (nstreams, n / 1e6, D, dec.taps.size),
            "metric": "input IQ Msamples/s", "value": nstreams * n / (ms * 1e-3) / 1e6, "ms_per_step": ms,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk, "peak_source": src,
                         "algorithmic_bytes": algo, "kernel": "fir_decimate_u8_tma<%d, 79, %d> (PIRIP_DECIM_NO_TMA=1: fir_decimate_u8_fast; PIRIP_DECIM_GENERIC=1: fir_decimate_kernel<true>)" % (D, 16 if D == 45 else 12)},
            "parity_vs_oracle": ok}


def bench_chain_c2(steps):
    """Config 2 end to end on the device: 1.8 MS/s u8 IQ -> fused csdr stage (/45, s16 as in [R README.md:109]) ->
    fsk_demod -c 2 40000 1000 --fsk_lower 500 -> bits; rate quoted per INPUT sample.  Two launches per step."""
    import orc
    import synth
    D, nstreams, nbits = 45, 592, 1500  # 1.5 s streams: 2.7 Msamples each, 3.2 GB of u8 in total
    base = []
    for c in range(8):
        x = synth.fsk_mod(synth.test_bits(nbits), 2, 1800000, 1000, 1000, 1000, timing_offset=37.0 * c, phase0=0.3 * c)
        x = synth.add_awgn(x, [None, 12.0, 9.0, 6.0][c % 4], 1800, 2, synth.rng(40 + c))
        base.append(synth.to_u8(x, amp=70.0))
    n = min(b.size for b in base) // 2
    base_t = torch.from_numpy(np.stack([b[:2 * n] for b in base])).cuda()
    iq = base_t[torch.arange(nstreams, device="cuda") % 8].contiguous()
    dec = pb.DecimPlan(D)
    nout = dec.out_count(n)
    mid = torch.empty((nstreams, 2 * nout), dtype=torch.int16, device="cuda")
    plan = pb.FskPlan(Fs=40000, Rs=1000, M=2, P=8, limits=(500, 20000), fmt=pb.FMT_S16)
    mf = plan.max_frames(nout)
    bits = torch.zeros((nstreams, mf * plan.nbits), dtype=torch.uint8, device="cuda")
    nf = torch.zeros(nstreams, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    sp = ctypes.c_void_p(s.cuda_stream)

    def fn():
        dec.run_dev(iq.data_ptr(), nstreams, n, iq.stride(0), None, mid.data_ptr(), nout, stream=sp)
        plan.demod_batch_dev(mid.data_ptr(), nstreams, nout, mid.stride(0) * 2, bits.data_ptr(), nf.data_ptr(), mf, stream=sp)

    ms = timed(fn, steps)
    _, ref_s = orc.csdr_decimate_stream(iq[5].cpu().numpy(), D)
    ref = orc.fsk_demod_stream(orc.demod_cfg(40000, 1000, 2, P=8, limits=(500, 20000), fmt=orc.FMT_S16), ref_s)
    k = int(nf[5])
    ok = k == ref["nframes"] and bool(np.array_equal(bits[5, :k * plan.nbits].cpu().numpy(), ref["bits"]))
    algo = nstreams * (2.0 * n + float(k * plan.nbits))
    pk, src = peak()
    ach = algo / (ms * 1e-3) / 1e9
    L = pb.lib()
    L.pirip_fsk_plan_kernel.restype = ctypes.c_char_p
    L.pirip_fsk_plan_kernel.argtypes = [ctypes.c_void_p]
    return {"workload": "chain C2: %d x %.2f Msample u8 IQ streams at 1.8 MS/s -> /45 (79 taps, s16) -> 2FSK Fs=40k Rs=1k --fsk_lower 500 -> bits"
                        % (nstreams, n / 1e6),
            "metric": "input IQ Msamples/s", "value": nstreams * n / (ms * 1e-3) / 1e6, "ms_per_step": ms, "launches_per_step": 2,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk, "peak_source": src,
                         "algorithmic_bytes": algo, "kernel": "fir_decimate_u8_tma<45, 79, 16> + " + L.pirip_fsk_plan_kernel(plan._h).decode()},
            "parity_vs_oracle": ok}


def bench_ldpc(steps):
    import orc
    l, v = pb.ldpc_code("H_256_512_4")
    N, K = 512, 256
    ncw = 65536
    g = np.random.default_rng(7)
    base = []
    for eb in (6.0, 3.0, 2.0, 1.0):
        ib = g.integers(0, 2, K, dtype=np.uint8)
        cw = np.concatenate([ib, pb.ldpc_encode(l, ib)]).astype(np.float64)
        sigma2 = 1.0 / (2 * 0.5 * 10 ** (eb / 10))
        y = (1 - 2 * cw) + g.standard_normal((ncw // 4, N)) * np.sqrt(sigma2)
        base.append((2 * y / sigma2).astype(np.float32))
    llr = torch.from_numpy(np.concatenate(base)).cuda()
    bits = torch.empty((ncw, N), dtype=torch.uint8, device="cuda")
    it = torch.empty(ncw, dtype=torch.int32, device="cuda")
    pcc = torch.empty(ncw, dtype=torch.int32, device="cuda")
    plan = pb.LdpcPlan(l)
    s = torch.cuda.current_stream()
    L = pb.lib()
    fn = lambda: L.pirip_ldpc_decode_batch_dev(plan._h, ctypes.c_void_p(llr.data_ptr()), ncw, ctypes.c_void_p(bits.data_ptr()),
                                               ctypes.c_void_p(it.data_ptr()), ctypes.c_void_p(pcc.data_ptr()),
                                               ctypes.c_void_p(s.cuda_stream))
    algo = ncw * (4.0 * N + N)
    pk, src = peak()
    out = []
    ref_bits = None
    for mode in ("sum_product", "sum_product_fast", "min_sum"):
        plan.set_mode(mode)
        ms = timed(fn, steps)
        ok = None
        if mode == "sum_product":
            idx = [0, ncw // 4 + 1, ncw // 2 + 2, 3 * ncw // 4 + 3]
            ok = True
            for i in idx:
                rb, rit, rpcc = orc.ldpc_decode(v, llr[i].cpu().numpy())
                ok = ok and np.array_equal(bits[i].cpu().numpy(), rb) and int(it[i]) == rit and int(pcc[i]) == rpcc
            ref_bits, ref_pcc = bits.clone(), pcc.clone()
            same = None
        else:
            both = (ref_pcc == 256) & (pcc == 256)
            same = bool(torch.equal(ref_bits[both][:, :K], bits[both][:, :K]))
        ach = algo / (ms * 1e-3) / 1e9
        out.append({"workload": "ldpc: %d codewords (512,256), Eb/N0 mix 6/3/2/1 dB, max 15 iterations" % ncw, "mode": mode,
                    "metric": "codewords/s", "value": ncw / (ms * 1e-3), "ms_per_step": ms, "mean_iterations": float(it.float().mean()),
                    "word_error_rate": float((pcc != 256).float().mean()),
                    "coded_Mbit_per_s": ncw * N / (ms * 1e-3) / 1e6,
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk, "peak_source": src,
                                 "algorithmic_bytes": algo, "kernel": "ldpc_decode_warp_kernel<%s>" % mode,
                                 "note": "compute-bound: message passing over ~1.5 k edges per iteration; LLRs in (4 B/bit), bits out (1 B/bit)"},
                    "parity_vs_oracle(bits,iter,pcc)": ok, "payload_equal_to_exact_mode_on_converged_words": same})
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.only in ("", "decim"):
        print(json.dumps(bench_decim(a.steps)))
    if a.only in ("", "decim30"):  # rtl_fsk -s 2400000 -a 80000 [R README.md:172]
        print(json.dumps(bench_decim(a.steps, D=30, n=2400000)))
    if a.only in ("", "chain"):
        print(json.dumps(bench_chain_c2(a.steps)))
    if a.only in ("", "ldpc"):
        for r in bench_ldpc(a.steps):
            print(json.dumps(r))
