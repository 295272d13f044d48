#!/usr/bin/env python
"""Secondary bench lines: measured at N = 1 beside the headline line of bench.py, same timing rules
(CUDA events on the launching stream, >= 3 warm-up steps, inputs larger than L2 or stated otherwise).

  chain_c4         BASELINE.json config 4: u8 captures of LDPC-coded bursts -> pirip_fsk_ldpc_rx_batch (demodulator with LLRs,
                   UW sync, LDPC, CRC on the device) -> rx_status + packed payload per modem frame; IQ Msamples/s, roofline
                   at 2 + 33/(544 * 24) = 2.0025 B/sample, the oracle chain on the host cores beside it
  c3_4fsk_llr      config 3: 4FSK Fs=240k Rs=10k, --fsk_lower 500 --fsk_upper 25000, LLRs out (2.3333 B/sample)
  decim_45 / chain_c2 / ldpc   the csdr stage alone (config 2's front end), config 2 end to end (decimate /45 + demodulate, two
                   launches) and the LDPC decoder in its three modes: tools/bench_extra.py
  c1_single_stream config 1: ONE 600 k-bit stream [R README.md:101-105] through tools/bin/fsk_demod (GPU, one frame per call as
                   the CLI loop does, and --batch = the whole stream in one launch) beside the CPU port on ONE thread

Stand-alone: python tools/bench_secondary.py [--only name] > profiles/r02_secondary.json
Test infrastructure imports (oracle) only in the CPU-baseline / parity legs.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if os.path.exists(p) else (6650.0, "fallback (B200_PROFILING.md)")


def timed(torch, fn, steps, warmup=3):
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


def threaded(fn, n, threads):
    ths = [threading.Thread(target=lambda t=t: [fn(i) for i in range(t, n, threads)]) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0


# ------------------------------------------------------------------ chain_c4
C4 = dict(Fs=240000, Rs=10000, M=2, P=8, limits=(500, 25000))
C4_CODE = "H_256_512_4"
C4_NSAMP = 60000  # 0.25-s captures: 50 modem frames, one burst of 3 coded frames (UW + 512 bits each) behind a preamble


def gen_coded_captures(torch, pb, ncap, seed, device, ebnos=(9.0, 7.0, 6.0, 5.0)):
    """[ncap, 2 * C4_NSAMP] u8: silence (noise only), 50-bit preamble, three FSK_LDPC frames with CRC (the transmitter's burst
    layout [R tx/rpitx_fsk.cpp:313-336,369-400]), silence; per-capture random gap, phase, timing, frequency offset.
    Returns (iq, list of the packed payloads sent per capture)."""
    import orc
    Fs, Rs = C4["Fs"], C4["Rs"]
    Ts = Fs // Rs
    nbits = C4_NSAMP // Ts
    modem = pb.FskLdpcModem(C4_CODE, M=2, Rs=Rs, Fs=Fs)
    g = np.random.default_rng(0x5049 + seed)
    pool, pool_bytes = [], []
    for k in range(64):
        payload = orc.payload_data_bits(modem.data_bits).copy()
        payload[8:40] = g.integers(0, 2, 32, dtype=np.uint8)
        payload = orc.insert_crc(payload)
        pool.append(modem.frame(payload))
        pool_bytes.append(np.packbits(payload).tobytes())
    modem.close()
    preamble = np.array([0, 1] * 25, np.uint8)
    bits = np.zeros((ncap, nbits), np.uint8)
    on = np.zeros((ncap, nbits), np.uint8)
    sent = []
    for c in range(ncap):
        gap = int(g.integers(200, 500))
        pick = g.integers(0, 64, 3)
        seq = np.concatenate([preamble] + [pool[i] for i in pick])
        bits[c, gap:gap + seq.size] = seq
        on[c, gap:gap + seq.size] = 1
        sent.append([pool_bytes[i] for i in pick])
    tg = torch.Generator(device=device)
    tg.manual_seed(0x50495250 + 77 + seed)
    bits_t = torch.from_numpy(bits).to(device)
    on_t = torch.from_numpy(on).to(device).to(torch.float64)
    out = torch.empty((ncap, 2 * C4_NSAMP), dtype=torch.uint8, device=device)
    n = torch.arange(C4_NSAMP, dtype=torch.float64, device=device)[None, :]
    chunk = 100
    for c0 in range(0, ncap, chunk):
        nc = min(chunk, ncap - c0)
        u = torch.rand((nc, 3), generator=tg, device=device, dtype=torch.float64)
        toff = u[:, 1:2] * Ts
        k = torch.floor((n - toff) / Ts).clamp_(min=0, max=nbits - 1).to(torch.int64)
        bit = torch.gather(bits_t[c0:c0 + nc], 1, k).to(torch.float64)
        a = torch.gather(on_t[c0:c0 + nc], 1, k)
        f = (10000.0 + (u[:, 2:3] * 2 - 1) * 300.0 + 10000.0 * bit) / Fs
        ph = u[:, 0:1] * 2 * np.pi + 2 * np.pi * torch.cumsum(f, dim=1)
        eb = torch.tensor([ebnos[(c0 + i) % len(ebnos)] for i in range(nc)], dtype=torch.float64, device=device)[:, None]
        sigma = torch.sqrt(Ts / (2.0 * 10 ** (eb / 10.0)))
        re = a * torch.cos(ph) + sigma * torch.randn((nc, C4_NSAMP), generator=tg, device=device, dtype=torch.float64)
        im = a * torch.sin(ph) + sigma * torch.randn((nc, C4_NSAMP), generator=tg, device=device, dtype=torch.float64)
        out[c0:c0 + nc, 0::2] = torch.clamp(torch.round(127.5 + 30.0 * re), 0, 255).to(torch.uint8)
        out[c0:c0 + nc, 1::2] = torch.clamp(torch.round(127.5 + 30.0 * im), 0, 255).to(torch.uint8)
    return out, sent


def oracle_chain(cfg, code, u8, Nbits):
    import orc
    ref = orc.fsk_demod_stream(cfg, u8)
    od = orc.Deframer(code, max_nbits=2 * Nbits)
    st, pl = [], []
    for k in range(ref["nframes"]):
        s, p = od.push(ref["llr"][k * Nbits:(k + 1) * Nbits])
        st.append(s)
        pl.append(np.packbits(p).tobytes() if s & orc.RX_BITS else None)
    od.close()
    return ref["nframes"], st, pl


def chain_c4(torch, pb, dev, cores, steps=5, ncap=5000, check=256, cpu=True):
    import orc
    _, code = pb.ldpc_code(C4_CODE)
    iq, sent = gen_coded_captures(torch, pb, ncap, 0, dev)
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    nsamp = C4_NSAMP
    per_mode = {}
    ref = None
    nchk = min(check, ncap)
    ocfg = orc.demod_cfg(C4["Fs"], C4["Rs"], C4["M"], P=C4["P"], limits=C4["limits"], soft=True)
    iq_h = iq[:nchk].cpu().numpy()
    algo_bps = 2.0 + 33.0 / (544 * 24)
    pk, src = peak()
    kernel_name = None
    for demod_mode, ldpc_mode in (("exact", "sum_product"), ("fast", "sum_product_fast"), ("fast", "min_sum")):
        note("chain_c4 %s + %s" % (demod_mode, ldpc_mode))
        plan = pb.RxPlan(C4_CODE, **C4, demod_mode=demod_mode, ldpc_mode=ldpc_mode, device=dev.index or 0)
        mf = plan.max_frames(nsamp)
        llr = torch.zeros((ncap, mf * plan.nbits), dtype=torch.float32, device=dev)
        nf = torch.zeros(ncap, dtype=torch.int32, device=dev)
        status = torch.zeros((ncap, mf), dtype=torch.uint8, device=dev)
        payload = torch.zeros((ncap, mf, plan.payload_bytes), dtype=torch.uint8, device=dev)
        fn = lambda: plan.rx_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), llr.data_ptr(), nf.data_ptr(), status.data_ptr(),
                                       payload.data_ptr(), mf, stream=sp)
        l0 = pb.kernel_launches()
        ms = timed(torch, fn, steps)
        launches = (pb.kernel_launches() - l0) // (steps + 3)
        # the demodulator alone, for the split
        fplan = pb.FskPlan(**C4, soft=True, device=dev.index or 0)
        fplan.set_mode(demod_mode)
        fsk_only = lambda: fplan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), None, nf.data_ptr(), mf, d_llr=llr.data_ptr(), stream=sp)
        ms_demod = timed(torch, fsk_only, steps)
        if kernel_name is None or demod_mode == "fast":
            kernel_name = fplan.kernel()
        fn()
        torch.cuda.synchronize()
        st_h, pl_h, nf_h = status.cpu().numpy(), payload.cpu().numpy(), nf.cpu().numpy()
        RXB = pb._lib.RX_BITS
        delivered = total = bogus = 0
        for c in range(ncap):
            ss = set(sent[c])
            total += len(sent[c])
            for k in np.nonzero(st_h[c, :nf_h[c]] & RXB)[0]:
                if pl_h[c, k].tobytes() in ss:
                    delivered += 1
                else:
                    bogus += 1
        entry = {"demod_mode": demod_mode, "ldpc_mode": ldpc_mode, "ms_per_step": ms, "ms_demodulator_alone": ms_demod,
                 "value": ncap * nsamp / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "launches_per_step": int(launches),
                 "packets_sent": total, "packets_delivered": delivered, "packets_not_among_the_sent": bogus,
                 "packet_error_rate": 1.0 - delivered / max(total, 1)}
        if demod_mode == "exact":
            # parity: status sequence and payload of every modem frame against the oracle chain (exact modes: must be identical)
            refs = [None] * nchk

            def work(c):
                refs[c] = oracle_chain(ocfg, code, iq_h[c], plan.nbits)
            threaded(work, nchk, cores)
            bad = 0
            for c in range(nchk):
                rnf, rst, rpl = refs[c]
                same = rnf == nf_h[c] and st_h[c, :rnf].tolist() == rst and all(
                    (pl_h[c, k].tobytes() == rpl[k]) if rst[k] & orc.RX_BITS else not pl_h[c, k].any() for k in range(rnf))
                bad += 0 if same else 1
            entry["parity"] = {"captures_checked": nchk, "captures_mismatching": bad,
                               "what": "rx_status sequence and payload bytes of every modem frame equal the oracle chain (orc.fsk_demod_stream -> orc.Deframer -> CRC)"}
            ref = refs
        else:
            # tolerance modes: the packets the oracle chain delivers on the checked captures against the ones delivered here
            o_del = g_del = both = 0
            for c in range(nchk):
                rnf, rst, rpl = ref[c]
                o = {rpl[k] for k in range(rnf) if rst[k] & orc.RX_BITS}
                gset = {pl_h[c, k].tobytes() for k in np.nonzero(st_h[c, :nf_h[c]] & RXB)[0]}
                o_del += len(o)
                g_del += len(gset)
                both += len(o & gset)
            entry["parity"] = {"captures_checked": nchk, "packets_oracle_chain": o_del, "packets_here": g_del, "packets_in_both": both,
                               "what": "tolerance modes: payloads are CRC-checked, so a delivered packet is right; the counts show what the tolerance costs"}
        per_mode[demod_mode + "+" + ldpc_mode] = entry
        plan.close()
        fplan.close()
        del llr, status, payload
    best = per_mode["fast+min_sum"] if per_mode["fast+min_sum"]["value"] > per_mode["fast+sum_product_fast"]["value"] else per_mode["fast+sum_product_fast"]
    ach = ncap * nsamp * algo_bps / (best["ms_per_step"] * 1e-3) / 1e9

    note("chain_c4 e2e")
    # e2e: host buffers through pirip_fsk_ldpc_rx_batch (H2D of the captures, D2H of status / payload inside the call)
    plan = pb.RxPlan(C4_CODE, **C4, demod_mode=best["demod_mode"], ldpc_mode=best["ldpc_mode"], device=dev.index or 0)
    h_iq = torch.empty((ncap, 2 * nsamp), dtype=torch.uint8, pin_memory=True)
    h_iq.copy_(iq)
    mf = plan.max_frames(nsamp)
    h_st = torch.empty((ncap, mf), dtype=torch.uint8, pin_memory=True)
    h_pl = torch.empty((ncap, mf, plan.payload_bytes), dtype=torch.uint8, pin_memory=True)
    h_nf = torch.empty(ncap, dtype=torch.int32, pin_memory=True)

    def e2e_step():
        rc = plan.L.pirip_fsk_ldpc_rx_batch(plan._h, ctypes.c_void_p(h_iq.data_ptr()), ncap, nsamp, ctypes.c_void_p(h_st.data_ptr()),
                                            ctypes.c_void_p(h_pl.data_ptr()), ctypes.c_void_p(h_nf.data_ptr()), None, mf)
        assert rc == 0
    for _ in range(2):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / steps
    plan.close()

    cpu_baseline = None
    if cpu:
        ncpu = min(ncap, 8 * cores)
        iq_c = iq[:ncpu].cpu().numpy()
        passes = 3
        threaded(lambda c: oracle_chain(ocfg, code, iq_c[c], 50), ncpu, cores)
        t = float(np.mean([threaded(lambda c: oracle_chain(ocfg, code, iq_c[c], 50), ncpu, cores) for _ in range(passes)]))
        cpu_baseline = {"value": ncpu * nsamp / t / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                        "sample": "%d of the %d captures x %d timed passes, %d threads; oracle chain (restatement of codec2 fsk_demod + "
                                  "FSK_LDPC deframer + sum-product LDPC + CRC)" % (ncpu, ncap, passes, cores)}
    return {"metric": "IQ Msamples/s decimate-free coded chain: fsk_demod (LLR) + UW sync + LDPC + CRC, 2-FSK Fs=240k Rs=10k, %s" % C4_CODE,
            "config": {"workload": "C4: %d x 0.25-s u8 IQ captures, each one burst (preamble + 3 FSK_LDPC frames of %s with CRC-16) in noise, "
                                   "Eb/N0 (channel bit) mix 9/7/6/5 dB, 2-FSK Fs=240k Rs=10k P=8, --fsk_lower 500 --fsk_upper 25000; "
                                   "rx_status byte + 32 packed payload bytes per modem frame out" % (ncap, C4_CODE),
                       "captures": ncap, "samples_per_capture": nsamp, "code_table": "stand-in for codec2's %s (DESIGN.md 5)" % C4_CODE,
                       "cache": "inputs (%.0f MB) larger than the 126 MB L2" % (ncap * nsamp * 2 / 1e6)},
            "value": best["value"], "unit": "Msamples/s", "ms_per_step": best["ms_per_step"], "modes": per_mode,
            "headline_mode": best["demod_mode"] + "+" + best["ldpc_mode"],
            "e2e": {"value": ncap * nsamp / e2e_s / 1e6, "unit": "Msamples/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": int(ncap * nsamp * 2),
                    "d2h_bytes_per_step": int(h_st.numel() + h_pl.numel() + 4 * ncap), "api": "pirip_fsk_ldpc_rx_batch (host buffers, pinned)"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk, "peak_source": src, "traffic": None,
                         "algorithmic_bytes_per_sample": algo_bps, "kernel": kernel_name + " + fsk_ldpc_rx_kernel (one warp per capture)",
                         "note": "two launches per step; both instruction-bound (DESIGN.md section 6)"},
            "cpu_baseline": cpu_baseline}


# ------------------------------------------------------------------ c3_4fsk_llr
def c3_4fsk_llr(torch, pb, dev, cores, steps=5, ncap=1250, seconds=1.0, check=64, cpu=True):
    import orc
    import bench
    Fs, Rs, M, P = 240000, 10000, 4, 8
    limits = (500, 25000)  # as BASELINE.json config 3 spells it; four tones 0.75 Rs apart fit the window
    nsamp = int(seconds * Fs)
    # the headline generator with four tones: symbols from bit pairs of the test frame
    g = torch.Generator(device=dev)
    g.manual_seed(0x50495250 + 333)
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
    plan = pb.FskPlan(Fs, Rs, M, P=P, limits=limits, soft=True, device=dev.index or 0)
    mf = plan.max_frames(nsamp)
    llr = torch.zeros((ncap, mf * plan.nbits), dtype=torch.float32, device=dev)
    bits = torch.zeros((ncap, mf * plan.nbits), dtype=torch.uint8, device=dev)
    nf = torch.zeros(ncap, dtype=torch.int32, device=dev)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lambda: plan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), bits.data_ptr(), nf.data_ptr(), mf, d_llr=llr.data_ptr(), stream=sp)
    ms = timed(torch, fn, steps)
    flagged, _ = plan.flagged(0)
    kernel_name = plan.kernel()
    nchk = min(check, ncap)
    ocfg = orc.demod_cfg(Fs, Rs, M, P=P, limits=limits, soft=True)
    iq_h, bits_h, llr_h, nf_h = iq[:nchk].cpu().numpy(), bits[:nchk].cpu().numpy(), llr[:nchk].cpu().numpy(), nf[:nchk].cpu().numpy()
    res = [None] * nchk

    def work(c):
        r = orc.fsk_demod_stream(ocfg, iq_h[c])
        k = r["nframes"] * plan.nbits
        same = r["nframes"] == nf_h[c] and np.array_equal(r["bits"], bits_h[c, :k])
        if same and k:
            gl, rl = llr_h[c, :k].reshape(r["nframes"], -1).astype(np.float64), r["llr"].reshape(r["nframes"], -1).astype(np.float64)
            d = np.abs(gl - rl) / np.maximum(np.abs(rl).max(axis=1, keepdims=True), 1e-30)
        else:
            d = np.array([0.0 if same else np.inf])
        res[c] = (same, float(d.max()))
    threaded(work, nchk, cores)
    algo_bps = 2.0 + 400.0 / 1200.0
    pk, src = peak()
    ach = ncap * nsamp * algo_bps / (ms * 1e-3) / 1e9
    cpu_baseline = None
    if cpu:
        ncpu = min(ncap, 2 * cores)
        iq_c = iq[:ncpu].cpu().numpy()
        threaded(lambda c: orc.fsk_demod_stream(ocfg, iq_c[c]), ncpu, cores)
        t = float(np.mean([threaded(lambda c: orc.fsk_demod_stream(ocfg, iq_c[c]), ncpu, cores) for _ in range(4)]))
        cpu_baseline = {"value": ncpu * nsamp / t / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                        "sample": "%d of the %d captures x 4 timed passes, %d threads; oracle restatement of codec2 fsk_demod -s" % (ncpu, ncap, cores)}
    return {"metric": "IQ Msamples/s fsk_demod 4-FSK Fs=240k Rs=10k, soft-decision LLR output",
            "config": {"workload": "C3: %d x %.1f-s u8 IQ captures, 4-FSK Fs=240k Rs=10k P=8 (tones 0.75 Rs apart from 937.5 Hz), --fsk_lower 500 "
                                   "--fsk_upper 25000, f32 LLRs + hard bits out" % (ncap, seconds), "captures": ncap, "samples_per_capture": nsamp,
                       "cache": "inputs (%.0f MB) larger than the 126 MB L2" % (ncap * nsamp * 2 / 1e6)},
            "value": ncap * nsamp / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms,
            "captures_left_to_the_exact_kernel_last_step": int(flagged),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk, "unit": "GB/s", "frac": ach / pk, "peak_source": src, "traffic": None,
                         "algorithmic_bytes_per_sample": algo_bps, "kernel": kernel_name},
            "parity": {"captures_checked": nchk, "bits_and_frames_equal_oracle": bool(all(r[0] for r in res)),
                       "captures_mismatching": int(sum(0 if r[0] else 1 for r in res)),
                       "max_llr_error_relative_to_the_frame_s_largest_llr": max(r[1] for r in res), "llr_tolerance": 1e-3},
            "cpu_baseline": cpu_baseline}


# ------------------------------------------------------------------ c1_single_stream
def c1_single_stream(pb, cores, nbits=600000):
    """fsk_get_test_bits - 600000 | modulator | fsk_demod -d -p 24 --fsk_lower 500 --fsk_upper 25000 2 240000 10000
    [R README.md:101-105]: ONE stream, a serial chain of frames.  GPU through the command-line tool (a launch per frame, and
    --batch: one launch), CPU port on one thread."""
    import orc
    import synth
    Fs, Rs = 240000, 10000
    bits = synth.test_bits(nbits)
    x = synth.fsk_mod(bits, 2, Fs, Rs, 10000, 10000, timing_offset=3.3, phase0=0.4)
    iq = synth.to_u8(synth.add_awgn(x, 9.0, Fs // Rs, 2, synth.rng(11)), amp=60.0)
    path = "/tmp/pirip_c1_stream.u8"
    iq.tofile(path)
    BIN = os.path.join(ROOT, "tools", "bin")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    base = [os.path.join(BIN, "fsk_demod"), "-d", "-p", "24", "--fsk_lower", "500", "--fsk_upper", "25000"]
    out = {}
    ref_bits = None
    cfg = orc.demod_cfg(Fs, Rs, 2, P=24, limits=(500, 25000))
    orc.fsk_demod_stream(cfg, iq[:240000])
    t0 = time.perf_counter()
    ref = orc.fsk_demod_stream(cfg, iq)
    t_cpu = time.perf_counter() - t0
    ref_bits = ref["bits"]
    for name, extra in (("gpu_cli_frame_per_call", []), ("gpu_cli_batch_one_launch", ["--batch"])):
        note("c1_single_stream " + name)
        t0 = time.perf_counter()
        r = subprocess.run(base + extra + ["2", str(Fs), str(Rs), path, "-"], capture_output=True, check=True, timeout=300)
        dt = time.perf_counter() - t0
        got = np.frombuffer(r.stdout, np.uint8)
        out[name] = {"wall_s_including_process_and_context_start": dt, "Msamples_per_s": iq.size / 2 / dt / 1e6,
                     "bits_equal_cpu_port": bool(got.size == ref_bits.size and np.array_equal(got, ref_bits))}
    ber, nb = synth.ber_vs_stream(ref_bits)
    return {"metric": "IQ Msamples/s, ONE 2-FSK stream Fs=240k Rs=10k P=24 (config 1: 600 k test bits)",
            "config": {"workload": "C1: one %.1f-s stream (%d test bits), Eb/N0 9 dB" % (iq.size / 2 / Fs, nbits), "samples": int(iq.size // 2)},
            "cpu_port_single_thread": {"Msamples_per_s": iq.size / 2 / t_cpu / 1e6, "wall_s": t_cpu, "cores": 1, "kind": "port", "ber": ber},
            "gpu": out,
            "note": "a single stream is one serial chain of frames (replicas only, DESIGN.md 7): the GPU's unit of parallelism is the capture, "
                    "so this line is latency, not throughput; the tool times include process start and CUDA context creation"}


def note(msg):
    print("[bench_secondary %.1f] %s" % (time.time() % 10000, msg), file=sys.stderr, flush=True)


def run_all(torch, pb, dev, cores, only="", cpu=True, **kw):
    res = {}
    def extra(which):
        # the csdr stage, the config-2 chain (decimate + demodulate, two launches) and the LDPC decoder: tools/bench_extra.py
        import bench_extra
        return {"decim_45": lambda: bench_extra.bench_decim(5), "chain_c2": lambda: bench_extra.bench_chain_c2(5),
                "ldpc": lambda: bench_extra.bench_ldpc(5)}[which]()

    for name, fn in (("chain_c4", lambda: chain_c4(torch, pb, dev, cores, cpu=cpu, **kw)),
                     ("c3_4fsk_llr", lambda: c3_4fsk_llr(torch, pb, dev, cores, cpu=cpu, **kw)),
                     ("c1_single_stream", lambda: c1_single_stream(pb, cores)),
                     ("decim_45", lambda: extra("decim_45")), ("chain_c2", lambda: extra("chain_c2")), ("ldpc", lambda: extra("ldpc"))):
        if only in ("", name):
            note(name + " ...")
            t0 = time.time()
            try:
                res[name] = fn()
            except Exception as e:  # one leg must not take the others with it
                res[name] = {"error": repr(e)}
            note("%s done in %.1f s" % (name, time.time() - t0))
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ncap", type=int, default=0, help="captures (default: the full workload of each leg)")
    a = ap.parse_args()
    import torch

    import pirip_b200 as pb
    os.environ.setdefault("PIRIP_LDPC_STANDIN_OK", "1")
    kw = {"ncap": a.ncap} if a.ncap else {}
    print(json.dumps(run_all(torch, pb, torch.device("cuda", 0), os.cpu_count() or 1, a.only, not a.no_cpu, **kw), indent=1))
