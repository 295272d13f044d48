#!/usr/bin/env python
"""bench.py — IQ Msamples/s through the FSK demodulator (BASELINE.json metric).

Workload (config 5 of BASELINE.json at the C1 modem settings): independent 1-second u8 IQ
captures, 2-FSK Fs=240 kHz Rs=10 kHz P=24 (fsk_demod --fsk_lower 500 --fsk_upper 25000 -d -p 24
2 240000 10000, [R README.md:105]), 1250 captures per GPU (10^4 over 8 GPUs), synthetic
fsk_get_test_bits bursts with random phase / timing / frequency / clock offsets and AWGN.
A "step" is one pass of the demodulator over the whole per-GPU batch.

  value  : whole-job IQ Msamples/s with the captures already resident in HBM (one launch of the
           fast demodulator kernel per step - exact tone estimates, guarded decisions, exact
           in-kernel replay of the frames whose decisions fall inside a margin; a capture it could
           not settle would be re-run by the exact kernel inside the same step - timed with CUDA
           events on the launching stream, max over ranks)
  e2e    : the same metric through the host-buffer C-ABI call pirip_fsk_demod_batch(): pinned
           host u8 in -> H2D -> kernel -> D2H bits, every step
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".
  secondary (N = 1): chain_c4 (config 4: demodulator + UW sync + LDPC + CRC on the device), c3_4fsk_llr (config 3),
           c1_single_stream (config 1 through the command-line tool, CPU port on one thread beside it)

`--impl reference` times the CPU implementation of the path (the oracle restatement of codec2's
fsk_demod, all host cores) on a bounded sample of the same workload.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FS, RS, M, P, NSYM = 240000, 10000, 2, 24, 50
F1, SPACING = 10000, 10000
LIMITS = (500, 25000)
TS = FS // RS
N_FRAME = TS * NSYM
BYTES_PER_SAMPLE_ALGO = 2.0 + NSYM / float(N_FRAME)  # u8 I+Q in, 1 byte per hard bit out

TEST_FRAME = "0101111110001000101010010010010100011000011111001110111111100111110011100011110001110001111000110101"


def gen_captures(torch, ncap, nsamp, seed, device, ebno_list=(None, 12.0, 9.0, 6.0)):
    """Synthetic u8 IQ captures [ncap, 2*nsamp]: CPFSK of the repeating
    fsk_get_test_bits frame, per-capture random phase, timing, frequency and clock offset, AWGN."""
    g = torch.Generator(device=device)
    g.manual_seed(0x50495250 + seed)
    frame = torch.tensor([int(c) for c in TEST_FRAME], dtype=torch.float64, device=device)
    out = torch.empty((ncap, 2 * nsamp), dtype=torch.uint8, device=device)
    chunk = 50
    n = torch.arange(nsamp, dtype=torch.float64, device=device)[None, :]
    for c0 in range(0, ncap, chunk):
        nc = min(chunk, ncap - c0)
        u = torch.rand((nc, 4), generator=g, device=device, dtype=torch.float64)
        phase0 = u[:, 0:1] * 2 * np.pi
        toff = u[:, 1:2] * TS
        foff = (u[:, 2:3] * 2 - 1) * 0.1 * RS
        ppm = (u[:, 3:4] * 2 - 1) * 50.0
        ts_eff = TS * (1.0 + ppm * 1e-6)
        k = torch.floor((n - toff) / ts_eff).clamp_(min=0).to(torch.int64)
        bit = frame[k % frame.numel()]
        f = (F1 + foff + SPACING * bit) / FS
        ph = phase0 + 2 * np.pi * torch.cumsum(f, dim=1)
        re, im = torch.cos(ph), torch.sin(ph)
        eb = torch.tensor([ebno_list[(c0 + i) % len(ebno_list)] or 1e9 for i in range(nc)], dtype=torch.float64,
                          device=device)[:, None]
        sigma = torch.sqrt(TS / (2.0 * 10 ** (eb / 10.0)))  # per component, unit-amplitude signal
        amp = torch.clamp(110.0 / (1.0 + 3.0 * sigma), max=100.0)
        re = re + sigma * torch.randn((nc, nsamp), generator=g, device=device, dtype=torch.float64)
        im = im + sigma * torch.randn((nc, nsamp), generator=g, device=device, dtype=torch.float64)
        out[c0:c0 + nc, 0::2] = torch.clamp(torch.round(127.5 + amp * re), 0, 255).to(torch.uint8)
        out[c0:c0 + nc, 1::2] = torch.clamp(torch.round(127.5 + amp * im), 0, 255).to(torch.uint8)
    return out


class ClockSampler:
    """nvidia-smi sampling of SM clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(iq_np, steps, warmup, threads):
    """Time the CPU oracle (restatement of codec2 fsk_demod) over the rows of iq_np with
    `threads` host threads (ctypes releases the GIL).  Returns (samples/s per step list)."""
    import orc
    cfg = orc.demod_cfg(FS, RS, M, P=P, Nsym=NSYM, limits=LIMITS)
    orc.lib()
    ncap = iq_np.shape[0]

    def work(rows):
        for r in rows:
            orc.fsk_demod_stream(cfg, iq_np[r])

    def one_pass():
        ths = [threading.Thread(target=work, args=(range(t, ncap, threads),)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_pass()
    return [one_pass() for _ in range(steps)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--captures", type=int, default=1250, help="captures per GPU (10^4 over 8 GPUs)")
    ap.add_argument("--seconds", type=float, default=1.0, help="capture length")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary lines (chain_c4, c3_4fsk_llr, c1_single_stream, decim_45, chain_c2, ldpc)")
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1
    # stdout carries exactly ONE line, the JSON: libraries that write to file descriptor 1 themselves (NCCL prints its
    # version banner there when NCCL_DEBUG is set) are sent to stderr, the JSON line goes to the original descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(line):
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nsamp = int(round(args.seconds * FS))
    cores = os.cpu_count() or 1
    config = {"workload": "C5/C1: %d x %.1f-s u8 IQ captures per GPU, 2-FSK Fs=240k Rs=10k P=24, "
                          "--fsk_lower 500 --fsk_upper 25000, hard bits out" % (args.captures, args.seconds),
              "captures_per_gpu": args.captures, "samples_per_capture": nsamp,
              "cache": "inputs (%.0f MB/GPU) larger than the 126 MB L2" % (args.captures * nsamp * 2 / 1e6),
              "ebno_db_mix": [None, 12.0, 9.0, 6.0]}

    import torch

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        ncap = min(args.captures, 16 * cores)
        iq = gen_captures(torch, ncap, nsamp, seed=0, device="cpu").numpy()
        times = cpu_reference_run(iq, args.steps, args.warmup, cores)
        t = float(np.mean(times))
        v = ncap * nsamp / t / 1e6
        line = {"impl": "reference", "metric": "IQ Msamples/s fsk_demod 2-FSK Fs=240k Rs=10k", "value": v,
                "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                                 "sample": "%d of the %d captures per step, one capture stream per thread; "
                                           "oracle restatement of codec2 fsk_demod (gcc -O2)" % (ncap, args.captures)},
                "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import pirip_b200 as pb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        # N > 1: every rank streams 600 MB per step from pinned host memory; keep the rank's threads - and with them the first
        # touch of its pinned buffers - on the CPUs next to its GPU (NVML's affinity mask), so that eight ranks do not pull their
        # captures across the socket interconnect.  N = 1 keeps every core for the CPU baseline.
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa = "rank threads bound to %d CPUs next to GPU %d" % (len(cpus), local_rank)
        except Exception as e:  # affinity is an optimisation, never a requirement
            numa = "not bound (%s)" % type(e).__name__
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    plan = pb.FskPlan(FS, RS, M, P=P, Nsym=NSYM, limits=LIMITS, fmt=pb.FMT_U8, device=local_rank)
    ncap = args.captures
    mf = plan.max_frames(nsamp)
    iq = gen_captures(torch, ncap, nsamp, seed=rank, device=dev)
    stride = iq.stride(0)
    assert iq.data_ptr() % 16 == 0 and stride % 16 == 0
    nbits_cap = mf * plan.nbits
    bits = torch.zeros((ncap, nbits_cap), dtype=torch.uint8, device=dev)
    bits_dense = bits
    nframes = torch.zeros(ncap, dtype=torch.int32, device=dev)
    # config 5: the one exchange on the path - gather the decoded bit streams.  The batch's bit array is packed 8 bits per
    # byte on the device first (pirip_pack_bits_dev, the flat [ncap * max_frames * Nbits] array padded to a byte); two buffers.
    total_bits = ncap * nbits_cap
    total_bits8 = (total_bits + 7) // 8 * 8
    if world > 1 and total_bits8 != total_bits:
        flat = torch.zeros(total_bits8, dtype=torch.uint8, device=dev)
        bits = flat[:total_bits].view(ncap, nbits_cap)
        bits_dense = bits
    nbytes = total_bits8 // 8
    packed = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    # The gather of the packed bit streams must not need SMs while the next step's kernel runs: the demodulator is persistent
    # with one CTA per SM and all of its shared memory, so an NCCL kernel on a side stream only gets to run between two launches
    # and, gated by the slowest rank, stalled every step at N >= 4 (9.5 ms per step against 3.8 ms, measured).  Each rank
    # therefore PUSHES its packed bits into every peer's gather buffer with copy-engine peer copies over NVLink (torch symmetric
    # memory: the peers' buffers mapped into this process), on a side stream beside the next kernel.  Where symmetric memory is
    # not available the NCCL all-gather runs on the compute stream, right after the packing kernel, when the SMs are free.
    peers, gathered, gather_kind = None, None, None
    if world > 1:
        try:
            if os.environ.get("PIRIP_BENCH_GATHER", "") == "nccl":
                raise RuntimeError("NCCL gather requested")
            import torch.distributed._symmetric_memory as symm
            gathered = [symm.empty((world, nbytes), dtype=torch.uint8, device=dev) for _ in range(2)]
            handles = [symm.rendezvous(t, dist.group.WORLD) for t in gathered]
            peers = [[h.get_buffer(p, (world, nbytes), torch.uint8) for p in range(world)] for h in handles]
            gather_kind = "peer copies over NVLink by the copy engines into every rank's gather buffer (torch symmetric memory), beside the next step's kernel"
        except Exception as e:
            peers = None
            gathered = [torch.empty((world, nbytes), dtype=torch.uint8, device=dev) for _ in range(2)]
            gather_kind = "NCCL all-gather of the packed bits on the compute stream (%s)" % (str(e)[:80] or type(e).__name__)
    done = [None, None]  # event: the pushes that read packed[k] have been issued and completed
    step_no = [0]

    def demod_into(bits_t):
        plan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, stride, bits_t.data_ptr(), nframes.data_ptr(), mf,
                             stream=ctypes.c_void_p(stream.cuda_stream))

    def step():
        demod_into(bits_dense)
        if world > 1:
            k = step_no[0] & 1
            step_no[0] += 1
            if done[k] is not None:
                stream.wait_event(done[k])  # packed[k] is about to be overwritten: its pushes of two steps ago must be through
            pb.pack_bits_dev(ctypes.c_void_p(bits_dense.data_ptr()), total_bits8, ctypes.c_void_p(packed[k].data_ptr()),
                             ctypes.c_void_p(stream.cuda_stream))
            if peers is not None:
                ev = torch.cuda.Event()
                ev.record(stream)
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    for p in range(world):
                        peers[k][(rank + p) % world][rank].copy_(packed[k], non_blocking=True)
                    done[k] = torch.cuda.Event()
                    done[k].record(side)
            else:
                dist.all_gather_into_tensor(gathered[k], packed[k])

    pending = []

    def sync_all():
        while pending:
            pending.pop(0).wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = pb.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    launches = pb.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    gather_ok = None
    if world > 1:
        # untimed: what landed in this rank's gather buffer against a plain NCCL all-gather of the same packed streams
        k_last = (step_no[0] - 1) & 1
        ref_g = torch.empty((world, nbytes), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(ref_g, packed[k_last])
        ok_t = torch.tensor([1 if torch.equal(ref_g, gathered[k_last]) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        gather_ok = bool(int(ok_t.item()))
        del ref_g
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    total_samples = world * ncap * nsamp
    value = total_samples / (ms_per_step * 1e-3) / 1e6

    # kernel-only duration for the roofline (N=1: the step IS the kernel launch)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record(stream)
    for _ in range(args.steps):
        demod_into(bits_dense)
    k1.record(stream)
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps
    flagged, _ = plan.flagged(0)
    replays_ok, replays_failed = plan.replays()

    # ---- e2e through the host-buffer C-ABI call: pinned host in, bits back, every step
    e2e = None
    if not args.no_e2e:
        h_iq = torch.empty((ncap, 2 * nsamp), dtype=torch.uint8, pin_memory=True)
        h_iq.copy_(iq)
        h_bits = torch.empty((ncap, nbits_cap), dtype=torch.uint8, pin_memory=True)
        h_nf = torch.empty(ncap, dtype=torch.int32, pin_memory=True)
        L = pb.lib()

        def e2e_step():
            rc = L.pirip_fsk_demod_batch(plan._h, ctypes.c_void_p(h_iq.data_ptr()), ncap, nsamp,
                                         ctypes.c_void_p(h_bits.data_ptr()), None, None, None,
                                         ctypes.c_void_p(h_nf.data_ptr()), mf)
            assert rc == 0, L.pirip_last_error()

        for _ in range(2):
            e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_s = float(dt.item()) / args.steps
        assert torch.equal(h_bits, bits_dense.cpu()), "e2e bits differ from device-resident bits"
        # the ceiling of this leg: the same bytes as one plain pinned-host -> device copy on every rank at once
        d_tmp = torch.empty_like(iq)
        for _ in range(2):
            d_tmp.copy_(h_iq, non_blocking=True)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d_tmp.copy_(h_iq, non_blocking=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d_s = float(dt.item()) / args.steps
        del d_tmp
        e2e = {"value": total_samples / e2e_s / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(ncap * nsamp * 2), "d2h_bytes_per_step": int(h_bits.numel() + 4 * ncap),
               "ms_per_step": e2e_s * 1e3, "api": "pirip_fsk_demod_batch (host buffers, pinned)",
               "h2d_ceiling": {"ms_per_step": h2d_s * 1e3, "GBps_per_gpu": ncap * nsamp * 2 / h2d_s / 1e9,
                               "Msamples_per_s": total_samples / h2d_s / 1e6,
                               "what": "one cudaMemcpyAsync of the step's captures from pinned host memory on every rank at once, nothing else"},
               "fraction_of_h2d_ceiling": h2d_s / e2e_s, "host_affinity": numa}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- parity spot check + BER (not timed): GPU bits vs oracle on a few captures
    import orc
    import synth
    cfg = orc.demod_cfg(FS, RS, M, P=P, Nsym=NSYM, limits=LIMITS)
    # every capture of rank 0 against the oracle (threaded, untimed): frame counts and hard bits must be identical
    nchk = ncap if not args.no_cpu else min(8, ncap)
    iq_h = iq[:nchk].cpu().numpy()
    bits_h = bits_dense[:nchk].cpu().numpy()
    nf_h = nframes[:nchk].cpu().numpy()
    mism = [0] * nchk
    bers = [0.0] * nchk

    def check(rows):
        for c in rows:
            ref = orc.fsk_demod_stream(cfg, iq_h[c])
            same = ref["nframes"] == nf_h[c] and np.array_equal(ref["bits"], bits_h[c, :nf_h[c] * plan.nbits])
            mism[c] = 0 if same else 1
            bers[c] = synth.ber_vs_stream(bits_h[c, :nf_h[c] * plan.nbits])[0]

    ths = [threading.Thread(target=check, args=(range(t, nchk, cores),)) for t in range(cores)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    ok = sum(mism) == 0
    ebmix = config["ebno_db_mix"]
    ber_by_ebno = {}
    for i, e in enumerate(ebmix):
        v = [bers[c] for c in range(i, nchk, len(ebmix))]
        theory = 0.0 if e is None else 0.5 * float(np.exp(-0.5 * 10 ** (e / 10.0)))
        ber_by_ebno[str(e)] = {"measured": float(np.mean(v)) if v else None, "noncoherent_2fsk_theory": theory, "captures": len(v)}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = ncap * nsamp * BYTES_PER_SAMPLE_ALGO / (kernel_ms * 1e-3) / 1e9
    Lk = pb.lib()
    Lk.pirip_fsk_plan_kernel.restype = ctypes.c_char_p
    Lk.pirip_fsk_plan_kernel.argtypes = [ctypes.c_void_p]
    kernel_name = Lk.pirip_fsk_plan_kernel(plan._h).decode()
    # DRAM bytes per launch from the committed ncu --set full capture of this same workload
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")  # written by tools/ncu_summary.py from the committed ncu capture
    if os.path.exists(tpath) and ncap == 1250 and nsamp == 240000:
        t = json.load(open(tpath))
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        ncpu = min(ncap, 4 * cores)
        cpu_passes = 12  # about a second of wall clock on 16 threads: 10-30 s of CPU work
        t = float(np.mean(cpu_reference_run(iq[:ncpu].cpu().numpy(), cpu_passes, 1, cores)))
        cpu_baseline = {"value": ncpu * nsamp / t / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                        "sample": "%d of the %d captures (%.1f Msamples) x %d timed passes, %d threads, one capture stream each; "
                                  "oracle restatement of codec2 fsk_demod (gcc -O2)" % (ncpu, ncap, ncpu * nsamp / 1e6, cpu_passes, cores)}

    line = {"metric": "IQ Msamples/s fsk_demod 2-FSK Fs=240k Rs=10k", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, **({"gather": gather_kind, "gather_equals_nccl_all_gather": gather_ok,
                                       "gather_bytes_per_rank_per_step": int(nbytes)} if world > 1 else {})), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes": ncap * nsamp * BYTES_PER_SAMPLE_ALGO, "peak_source": peak_src,
                         "kernel": kernel_name,
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE_ALGO,
                         "note": "bound by instruction issue (FFT butterflies, mixing, integration), not HBM: DESIGN.md section 6"},
            "fast_mode": {"mode": plan.mode, "captures_left_to_the_exact_kernel_last_step": int(flagged),
                          "frames_replayed_exactly_in_kernel_total": int(replays_ok), "replays_failed_total": int(replays_failed),
                          "ppm_threshold_near_ties_total": plan.ppm_near_ties()},
            "cpu_baseline": cpu_baseline,
            "parity": {"captures_checked": nchk, "captures_per_gpu": ncap, "captures_mismatching": int(sum(mism)),
                       "bits_and_frames_equal_oracle": bool(ok), "ber_by_ebno_db": ber_by_ebno}}
    # secondary lines at N = 1: the coded chain of config 4, config 3 with LLR output, config 1 as ONE
    # stream through the command-line tool; each with its own roofline / CPU-port figure (tools/bench_secondary.py)
    if world == 1 and not args.no_secondary and ncap == 1250 and nsamp == 240000:
        del iq, bits, bits_dense
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_secondary
        try:
            line["secondary"] = bench_secondary.run_all(torch, pb, dev, cores, cpu=not args.no_cpu)
        except Exception as e:  # the headline line must not be lost to a secondary leg
            line["secondary"] = {"error": repr(e)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
