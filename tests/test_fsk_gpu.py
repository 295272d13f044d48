"""GPU parity: the CUDA demodulator (through the C-ABI) against the CPU oracle.

Bar (task statement / SURVEY.md 8d "BER match"): hard bits, frame counts and nin decisions
bit-exact.  This implementation goes further: the kernel evaluates the same IEEE-754
operation sequence as the oracle, so every float it reports (f_est, timing, SNR, rx_filt,
LLRs) is required to be BIT-IDENTICAL too (tolerance 0, compared as uint32 patterns).
"""
import zlib

import numpy as np
import pytest

import orc
import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_mode")]

# name -> (plan kwargs, generator kwargs)
CONFIGS = {
    # C1: [R README.md:105] fsk_demod --fsk_lower 500 --fsk_upper 25000 -d -p 24 2 240000 10000
    "c1_2fsk_240k_10k_p24_limits": (dict(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000)),
                                    dict(f1=10000, spacing=10000)),
    # C1 variant: [R test/loopback_rtl_sdr.sh:16] no limits, rx tuned 2*Rb low, shift 2*Rs
    "c1_2fsk_240k_10k_p24_nolimits": (dict(Fs=240000, Rs=10000, M=2, P=24), dict(f1=20000, spacing=20000)),
    # C3 sane variant: 4FSK, default P=8, soft decisions
    "c3_4fsk_240k_10k_soft": (dict(Fs=240000, Rs=10000, M=4, P=8, limits=(500, 45000), soft=True),
                              dict(f1=10000, spacing=10000)),
    # C3 as literally specified: limits 500..25000, 0.75*Rs tone spacing
    "c3_4fsk_literal": (dict(Fs=240000, Rs=10000, M=4, P=8, limits=(500, 25000), soft=True),
                        dict(f1=937.5, spacing=7500)),
    # mask estimator [R README.md:262]: rtl_fsk -a 200000 -r 10000 -m 4 --mask 10000
    "mask_4fsk_200k_10k": (dict(Fs=200000, Rs=10000, M=4, P=20, tone_spacing=10000, mask=True, soft=True),
                           dict(f1=10000, spacing=10000)),
    # C2 modem side: 2FSK Fs=40k Rs=1k default P, --fsk_lower 500 [R README.md:109]
    "c2_2fsk_40k_1k": (dict(Fs=40000, Rs=1000, M=2, P=8, limits=(500, 20000)), dict(f1=1000, spacing=1000)),
}


def make_batch(pk, gk, ncap, nbits, ebnos, seed0):
    caps = []
    for c in range(ncap):
        u8, _ = synth.capture_u8(seed0 + c, nbits, pk["M"], pk["Fs"], pk["Rs"], gk["f1"], gk["spacing"],
                                 ebno_db=ebnos[c % len(ebnos)])
        caps.append(u8)
    n = min(len(x) for x in caps)
    return np.stack([x[:n] for x in caps])


def oracle_cfg(pk):
    return orc.demod_cfg(pk["Fs"], pk["Rs"], pk["M"], P=pk.get("P", 8), Nsym=pk.get("Nsym", 50),
                         f1_tx=pk.get("f1_tx", 1200), tone_spacing=pk.get("tone_spacing", 400),
                         limits=pk.get("limits"), mask=pk.get("mask", False), fmt=pk.get("fmt", orc.FMT_U8),
                         soft=pk.get("soft", False))


def bits_equal(a, b):
    return a.view(np.uint32 if a.dtype == np.float32 else a.dtype).tolist() == \
        b.view(np.uint32 if b.dtype == np.float32 else b.dtype).tolist()


def compare(plan, pk, iq, out):
    cfg = oracle_cfg(pk)
    for c in range(iq.shape[0]):
        ref = orc.fsk_demod_stream(cfg, iq[c])
        nf = ref["nframes"]
        assert out["nframes"][c] == nf, f"capture {c}: frame count {out['nframes'][c]} != {nf}"
        got_bits = out["bits"][c, :nf * plan.nbits]
        nbad = int((got_bits != ref["bits"]).sum())
        assert nbad == 0, f"capture {c}: {nbad} hard-bit mismatches of {got_bits.size}"
        st = out["stats"][c, :nf]
        for name in st.dtype.names:
            a, b = np.ascontiguousarray(st[name]), np.ascontiguousarray(ref["stats"][name])
            if not bits_equal(a, b):
                bad = np.argwhere(a.view(np.uint32).reshape(a.shape) != b.view(np.uint32).reshape(b.shape))
                f0 = bad[0][0]
                raise AssertionError(f"capture {c}: stats.{name} differs first at frame {f0}: gpu {a[f0]} oracle {b[f0]}")
        assert bits_equal(out["rx_filt"][c, :nf * plan.nfilt], ref["rx_filt"]), f"capture {c}: rx_filt differs"
        if pk.get("soft"):
            assert bits_equal(out["llr"][c, :nf * plan.nbits], ref["llr"]), f"capture {c}: LLRs differ"


@pytest.mark.parametrize("name", list(CONFIGS))
def test_batch_matches_oracle(name):
    import pirip_b200 as pb
    pk, gk = CONFIGS[name]
    plan = pb.FskPlan(**pk)
    iq = make_batch(pk, gk, ncap=6, nbits=3000 * (2 if pk["M"] == 4 else 1) if pk["Rs"] > 1000 else 600,
                    ebnos=[None, 12.0, 9.0, 6.0, 3.0, 0.0], seed0=zlib.crc32(name.encode()) % 1000)
    out = plan.demod_batch(iq, want_llr=pk.get("soft", False), want_rx_filt=True, want_stats=True)
    compare(plan, pk, iq, out)
    plan.close()


def test_many_captures_per_sm_lockstep():
    """The batch kernel (fsk_kernel2.cu) demodulates several captures per CTA in lockstep: captures with
    different noise, timing and clock offsets (so their nin decisions differ inside one CTA) must come out exactly
    as the per-capture kernel (fsk_kernel.cu, PIRIP_FSK_K2=0) and the oracle produce them."""
    import os
    import torch
    import pirip_b200 as pb
    pk, gk = CONFIGS["c1_2fsk_240k_10k_p24_limits"]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    ncap = 3 * sms + 7  # 3..4 captures per CTA, uneven
    base = [synth.capture_u8(7000 + c, 2500, 2, pk["Fs"], pk["Rs"], gk["f1"], gk["spacing"],
                             ebno_db=[None, 12.0, 9.0, 6.0, 3.0][c % 5])[0] for c in range(24)]
    n = min(len(x) for x in base)
    base = np.stack([x[:n] for x in base])
    # distinct streams per capture: rotate the 24 base captures by a capture-dependent number of samples
    iq = np.stack([np.roll(base[c % 24], -2 * ((7 * c) % 977)) for c in range(ncap)])
    old = os.environ.get("PIRIP_FSK_K2")
    try:
        os.environ["PIRIP_FSK_K2"] = "0"
        p_ref = pb.FskPlan(**pk)
        ref = p_ref.demod_batch(iq, want_rx_filt=True, want_stats=True)
        p_ref.close()
        os.environ["PIRIP_FSK_K2"] = "1"
        plan = pb.FskPlan(**pk)
        out = plan.demod_batch(iq, want_rx_filt=True, want_stats=True)
    finally:
        if old is None:
            os.environ.pop("PIRIP_FSK_K2", None)
        else:
            os.environ["PIRIP_FSK_K2"] = old
    assert np.array_equal(out["nframes"], ref["nframes"])
    assert np.array_equal(out["bits"], ref["bits"])
    assert out["stats"].tobytes() == ref["stats"].tobytes(), "per-frame floats differ from the per-capture kernel"
    assert out["rx_filt"].tobytes() == ref["rx_filt"].tobytes()
    # some captures must have slipped (nin != N) for the test to mean anything
    nin = out["stats"]["nin_next"]
    assert (nin[out["stats"]["sample_offset"] > 0] != 1200).any()
    sub = list(range(0, ncap, max(1, ncap // 10)))
    compare(plan, pk, iq[sub], {k: (v[sub] if isinstance(v, np.ndarray) else v) for k, v in out.items()})
    plan.close()


@pytest.mark.parametrize("name", ["c1_2fsk_240k_10k_p24_limits", "c3_4fsk_240k_10k_soft"])
def test_cuda_exact_path_against_the_glibc_build_of_the_oracle(name):
    """The oracle and the CUDA kernels share pirip_b200/csrc/exact_math.h, so their bit-for-bit agreement says nothing about those
    sin / cos / atan2 / exp / log bodies.  Here the CUDA path meets the oracle built with glibc's own functions
    (liboracle_libm.so, what upstream codec2 links): hard bits, frame counts, nin and tone estimates identical, every float
    within 1e-4 (norm_rx_timing, absolute), 5e-4 (rx_filt) and 1e-3 (LLR, both relative to the frame's largest value)."""
    import pirip_b200 as pb
    pk, gk = CONFIGS[name]
    iq = make_batch(pk, gk, 6, 1500 if pk["M"] == 2 else 3000, [None, 9.0, 6.0], seed0=4100)
    plan = pb.FskPlan(**pk)
    out = plan.demod_batch(iq, want_llr=pk.get("soft", False), want_rx_filt=True, want_stats=True)
    plan.close()
    M = pk["M"]
    for c in range(iq.shape[0]):
        ref = orc.fsk_demod_stream(oracle_cfg(pk), iq[c], libm=True)
        nf = int(out["nframes"][c])
        assert nf == ref["nframes"]
        nb = ref["bits"].size // nf
        st = out["stats"][c, :nf]
        assert np.array_equal(st["nin_next"], ref["stats"]["nin_next"])
        assert np.array_equal(st["f_est"], ref["stats"]["f_est"])
        differ = out["bits"][c, :nf * nb] != ref["bits"]
        assert differ.mean() < 2e-4  # a decision can only flip where the two libms round an energy tie differently
        d = np.abs(st["norm_rx_timing"].astype(np.float64) - ref["stats"]["norm_rx_timing"])
        assert np.minimum(d, 1.0 - d).max() < 1e-4
        gf = out["rx_filt"][c, :nf * M * 50].reshape(nf, -1).astype(np.float64)
        rf = ref["rx_filt"].reshape(nf, -1).astype(np.float64)
        assert (np.abs(gf - rf) / np.maximum(rf.max(axis=1, keepdims=True), 1e-30)).max() < 5e-4  # 4FSK at 6 dB: 1.5e-4 between the two oracle builds
        if pk.get("soft"):
            gl = out["llr"][c, :nf * nb].reshape(nf, -1).astype(np.float64)
            rl = ref["llr"].reshape(nf, -1).astype(np.float64)
            assert (np.abs(gl - rl) / np.maximum(np.abs(rl).max(axis=1, keepdims=True), 1e-30)).max() < 1e-3
