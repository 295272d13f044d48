#!/usr/bin/env python
"""fast_probe.py — how far is the tolerance-mode kernel (fsk_fast.cu) from the oracle?

Runs the fast kernel WITHOUT the exact re-run (PIRIP_FSK_FAST_NO_RERUN / NO_EXIT: diagnostics) on synthetic batches
of the test configurations, runs the CPU oracle on the same bytes (threaded), and reports per configuration:
  - discrete outputs: frame counts, nin_next, f_est / f2_est (expected identical), hard-bit mismatches
  - float outputs: max |d norm_rx_timing|, rx_filt error relative to the frame's largest magnitude
  - the DECISION-MARGIN error: (e_best - e_second)/(e_best + e_second) fast vs oracle over every symbol - what the
    symbol guard has to cover - and | |norm| - 0.25 | error for the nin guard
  - which mismatching bits / nin decisions were NOT covered by the flags (must be zero)
Writes gpurun_out/fast_probe.json.  Test infrastructure (imports the oracle).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

os.environ.setdefault("PIRIP_FSK_FAST_NO_RERUN", "1")
os.environ.setdefault("PIRIP_FSK_FAST_NO_EXIT", "1")

import orc  # noqa: E402
import pirip_b200 as pb  # noqa: E402
import synth  # noqa: E402
from test_fsk_gpu import CONFIGS, make_batch, oracle_cfg  # noqa: E402


def oracle_batch(cfg, iq, threads):
    out = [None] * iq.shape[0]

    def work(rows):
        for r in rows:
            out[r] = orc.fsk_demod_stream(cfg, iq[r])

    ths = [threading.Thread(target=work, args=(range(t, iq.shape[0], threads),)) for t in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out


def margins(filt, M, Nsym):
    """(best - second)/(best + second) of the tone energies, per symbol; filt: [nframes, M*Nsym] tone-major"""
    e = filt.reshape(-1, M, Nsym).astype(np.float64) ** 2
    s = np.sort(e, axis=1)
    best, second = s[:, -1, :], s[:, -2, :]
    return (best - second) / np.maximum(best + second, 1e-300)


def probe(name, pk, gk, ncap, nbits, ebnos, threads):
    iq = make_batch(pk, gk, ncap, nbits, ebnos, seed0=7000)
    plan = pb.FskPlan(**pk, device=0)
    assert plan.mode == "fast", plan.kernel()
    t0 = time.time()
    out = plan.demod_batch(iq, want_llr=bool(pk.get("soft")), want_rx_filt=True, want_stats=True)
    nflag, flags = plan.flagged(ncap)
    rep = plan.replays()
    t_gpu = time.time() - t0
    refs = oracle_batch(oracle_cfg(pk), iq, threads)
    M, Nsym = pk["M"], pk.get("Nsym", 50)
    r = dict(config=name, captures=ncap, frames=0, symbols=0, flagged=int(nflag), nframes_mismatch=0, nin_mismatch=0, f_est_mismatch=0,
             bit_mismatch=0, captures_with_bit_mismatch=0, unflagged_bad_captures=0, max_dnorm=0.0, max_filt_rel=0.0,
             max_margin_err=0.0, max_llr_rel=0.0, max_dsnr_rel=0.0)
    merrs, dn_all, worst, rho_all = [], [], [], []
    for c in range(ncap):
        ref = refs[c]
        nf = min(ref["nframes"], int(out["nframes"][c]))
        bad = False
        if ref["nframes"] != out["nframes"][c]:
            r["nframes_mismatch"] += 1
            bad = True
        st, rs = out["stats"][c, :nf], ref["stats"][:nf]
        # frames are comparable only while the discrete state agrees
        same = (st["nin_next"] == rs["nin_next"]) & (st["sample_offset"] == rs["sample_offset"])
        ok_upto = nf if same.all() else int(np.argmin(same)) + 1
        if not same.all():
            r["nin_mismatch"] += 1
            bad = True
        fe = (st["f_est"][:ok_upto].view(np.uint32) != rs["f_est"][:ok_upto].view(np.uint32)).any() or \
             (st["f2_est"][:ok_upto].view(np.uint32) != rs["f2_est"][:ok_upto].view(np.uint32)).any()
        if fe:
            r["f_est_mismatch"] += 1
            bad = True
        nb = plan.nbits
        gb, rb = out["bits"][c, :ok_upto * nb], ref["bits"][:ok_upto * nb]
        nbad = int((gb != rb).sum())
        r["bit_mismatch"] += nbad
        if nbad:
            r["captures_with_bit_mismatch"] += 1
            bad = True
        if bad and flags[c] == 0:
            r["unflagged_bad_captures"] += 1
        r["frames"] += ok_upto
        r["symbols"] += ok_upto * Nsym
        dn = np.abs(st["norm_rx_timing"][:ok_upto].astype(np.float64) - rs["norm_rx_timing"][:ok_upto])
        dn = np.minimum(dn, 1.0 - dn)  # wrap at +-0.5
        dn_all.append(dn)
        if os.environ.get("PIRIP_FSK_FAST_DIAG"):
            rho = st["foff"][:ok_upto].astype(np.float64)
            rho_all.append(rho)
        gf = out["rx_filt"][c, :ok_upto * plan.nfilt].reshape(ok_upto, -1).astype(np.float64)
        rf = ref["rx_filt"][:ok_upto * plan.nfilt].reshape(ok_upto, -1).astype(np.float64)
        scale = np.maximum(rf.max(axis=1, keepdims=True), 1e-30)
        r["max_filt_rel"] = max(r["max_filt_rel"], float((np.abs(gf - rf) / scale).max()) if ok_upto else 0.0)
        me_c = np.abs(margins(gf, M, Nsym) - margins(rf, M, Nsym))
        merrs.append(me_c.ravel())
        if ok_upto:
            fw = int(np.argmax(me_c.max(axis=1)))
            worst.append((float(me_c[fw].max()), c, ebnos[c % len(ebnos)], fw, float(dn[fw]), float(rs["norm_rx_timing"][fw]),
                          float(rs["SNRest"][fw])))
        for k in ("SNRest", "v_est", "rx_sig_pow"):
            a, b = st[k][:ok_upto].astype(np.float64), rs[k][:ok_upto].astype(np.float64)
            r["max_dsnr_rel"] = max(r["max_dsnr_rel"], float((np.abs(a - b) / np.maximum(np.abs(b), 1e-30)).max()) if ok_upto else 0.0)
        if pk.get("soft"):
            gl = out["llr"][c, :ok_upto * nb].reshape(ok_upto, -1).astype(np.float64)
            rl = ref["llr"][:ok_upto * nb].reshape(ok_upto, -1).astype(np.float64)
            sc = np.maximum(np.abs(rl).max(axis=1, keepdims=True), 1e-30)
            r["max_llr_rel"] = max(r["max_llr_rel"], float((np.abs(gl - rl) / sc).max()) if ok_upto else 0.0)
    me = np.concatenate(merrs) if merrs else np.zeros(1)
    dn = np.concatenate(dn_all) if dn_all else np.zeros(1)
    r["max_margin_err"] = float(me.max())
    r["margin_err_pct"] = {p: float(np.percentile(me, p)) for p in (50, 99, 99.99)}
    r["max_dnorm"] = float(dn.max())
    r["dnorm_pct"] = {p: float(np.percentile(dn, p)) for p in (50, 99, 99.99)}
    if rho_all:
        rho = np.concatenate(rho_all)
        k = dn * rho
        r["rho_pct"] = {p: float(np.percentile(rho, p)) for p in (0.01, 1, 50)}
        r["dnorm_x_rho_pct"] = {p: float(np.percentile(k, p)) for p in (50, 99, 99.99, 100)}
        # per rho-decade: max dnorm
        r["dnorm_by_rho"] = [(float(lo), int(((rho >= lo) & (rho < lo * 10)).sum()), float(dn[(rho >= lo) & (rho < lo * 10)].max()) if ((rho >= lo) & (rho < lo * 10)).any() else 0.0)
                             for lo in (1e-4, 1e-3, 1e-2, 1e-1)]
    r["flag_reasons"] = {k: int(((flags & v) != 0).sum()) for k, v in (("sym", 1), ("tim", 2), ("ppm", 4), ("nan", 8))}
    r["flag_words"] = [(c, ebnos[c % len(ebnos)], int(flags[c] & 0xff), int((flags[c] >> 8) & 0xfff) - 1, int(flags[c] >> 20)) for c in range(ncap) if flags[c]]
    worst.sort(reverse=True)
    r["worst_margin_frames(err,cap,ebno,frame,dnorm,norm,SNRest)"] = worst[:6]
    r["replays(ok,failed)"] = rep
    r["kernel"] = plan.kernel()
    r["gpu_call_s"] = t_gpu
    plan.close()
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--captures", type=int, default=64)
    ap.add_argument("--nbits", type=int, default=2500)
    ap.add_argument("--configs", default="all")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fast_probe.json"))
    args = ap.parse_args()
    threads = os.cpu_count() or 4
    names = list(CONFIGS) if args.configs == "all" else args.configs.split(",")
    res = []
    for n in names:
        pk, gk = CONFIGS[n]
        r = probe(n, pk, gk, args.captures, args.nbits, [None, 12.0, 9.0, 6.0, 4.0], threads)
        print(json.dumps(r), flush=True)
        res.append(r)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
