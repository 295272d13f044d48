"""The oracle against an independent float64 model written from SURVEY.md Appendix A (tests/np_model.py) and against theory:
a check of the oracle's transcription that does not share a line with oracle/orc_fsk.c.  CPU only."""
import numpy as np
import pytest

import np_model
import orc
import synth

CASES = {
    # the golden configurations (tests/test_golden.py / tests/test_fsk_gpu.py)
    "c1": (dict(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000)), dict(f1=10000, spacing=10000)),
    "c1_nolimits": (dict(Fs=240000, Rs=10000, M=2, P=24), dict(f1=20000, spacing=20000)),
    "c3": (dict(Fs=240000, Rs=10000, M=4, P=8, limits=(500, 45000)), dict(f1=10000, spacing=10000)),
    "mask": (dict(Fs=200000, Rs=10000, M=4, P=20, tone_spacing=10000, mask=True), dict(f1=10000, spacing=10000)),
    "c2": (dict(Fs=40000, Rs=1000, M=2, P=8, limits=(500, 20000)), dict(f1=1000, spacing=1000)),
    "negative_window": (dict(Fs=80000, Rs=10000, M=2, P=8, limits=(-40000, 40000)), dict(f1=-20000, spacing=10000)),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("ebno", [None, 9.0, 6.0])
def test_oracle_agrees_with_the_independent_model(name, ebno):
    pk, gk = CASES[name]
    nbits = 1200 if pk["Rs"] >= 10000 else 700
    u8, _ = synth.capture_u8(60 + len(name), nbits * (1 if pk["M"] == 2 else 2), pk["M"], pk["Fs"], pk["Rs"], gk["f1"], gk["spacing"], ebno_db=ebno)
    cfg = orc.demod_cfg(pk["Fs"], pk["Rs"], pk["M"], P=pk["P"], tone_spacing=pk.get("tone_spacing", 400), limits=pk.get("limits"),
                        mask=pk.get("mask", False))
    ref = orc.fsk_demod_stream(cfg, u8)
    x = ((u8.astype(np.float64) - 127.0) / 128.0).view(np.complex128) if False else \
        ((u8[0::2].astype(np.float64) - 127.0) / 128.0) + 1j * ((u8[1::2].astype(np.float64) - 127.0) / 128.0)
    mdl = np_model.Model(pk["Fs"], pk["Rs"], pk["M"], P=pk["P"], tone_spacing=pk.get("tone_spacing", 400), limits=pk.get("limits"),
                         mask=pk.get("mask", False))
    out = np_model.demod_stream(mdl, x)
    M, Nsym, nb = pk["M"], 50, mdl.Nbits
    assert len(out) == ref["nframes"] and ref["nframes"] >= 10
    st = ref["stats"]
    checked = skipped = 0
    for k, fr in enumerate(out):
        # discrete outputs: the tone estimates (FFT bins) and the nin sequence are identical
        assert np.allclose(st["f_est"][k][:M], fr["f_est"], rtol=0, atol=1e-3), (k, st["f_est"][k], fr["f_est"])
        assert np.allclose(st["f2_est"][k][:M], fr["f2_est"], rtol=0, atol=1e-3), k
        if fr["timing_margin"] > 1e-4:
            assert int(st["nin_next"][k]) == fr["nin_next"], k
        # float outputs within float32 rounding of the double-precision model (1e-4)
        d = abs(float(st["norm_rx_timing"][k]) - fr["norm_rx_timing"])
        assert min(d, 1.0 - d) < 1e-4, (k, d)
        rf = ref["rx_filt"][k * M * Nsym:(k + 1) * M * Nsym].reshape(M, Nsym).astype(np.float64)
        assert np.abs(rf - fr["rx_filt"]).max() <= 1e-4 * max(fr["rx_filt"].max(), 1e-9), k
        # hard bits identical wherever the decision is not a float32-rounding coin flip
        safe = np.repeat(fr["margin"] > 1e-4, 1 if M == 2 else 2)
        assert np.array_equal(ref["bits"][k * nb:(k + 1) * nb][safe], fr["bits"][safe]), k
        checked += int(safe.sum())
        skipped += int((~safe).sum())
    assert skipped <= 0.002 * (checked + skipped) + 2


def _ber(M, ebno_db, seed, nbits=60000):
    Fs, Rs, P = 80000, 10000, 8
    g = synth.rng(seed)
    bits = g.integers(0, 2, nbits, dtype=np.uint8)
    f1, spacing = 5000, 10000
    x = synth.fsk_mod(bits, M, Fs, Rs, f1, spacing, timing_offset=3.0, phase0=0.5)
    u8 = synth.to_u8(synth.add_awgn(x, ebno_db, Fs // Rs, M, g), amp=35.0)
    ref = orc.fsk_demod_stream(orc.demod_cfg(Fs, Rs, M, P=P, limits=(500, 39000)), u8)
    rx = ref["bits"]
    # align: the demodulator's first symbol boundary is a fixed delay; best of a few shifts
    best = 1.0
    skip = 400 * (1 if M == 2 else 2)
    for sh in range(0, 8):
        n = min(rx.size - skip, bits.size - skip - sh)
        for sgn in (0,):
            e = float((rx[skip:skip + n] != bits[skip + sh:skip + sh + n]).mean())
            best = min(best, e)
        n2 = min(rx.size - skip - sh, bits.size - skip)
        best = min(best, float((rx[skip + sh:skip + sh + n2] != bits[skip:skip + n2]).mean()))
    return best


@pytest.mark.parametrize("ebno_db", [6.0, 8.0])
def test_4fsk_ber_against_noncoherent_theory(ebno_db):
    """Non-coherent orthogonal M-FSK: Ps = sum_{k=1}^{M-1} (-1)^(k+1) C(M-1,k)/(k+1) exp(-k/(k+1) Es/N0), Pb = Ps M/(2(M-1))
    (today only 2FSK was checked, tests/test_oracle_known_answers.py)."""
    from math import comb, exp, log2
    M = 4
    esno = 10 ** (ebno_db / 10) * log2(M)
    ps = sum((-1) ** (k + 1) * comb(M - 1, k) / (k + 1) * exp(-k / (k + 1) * esno) for k in range(1, M))
    pb = ps * M / (2 * (M - 1))
    got = _ber(M, ebno_db, 17)
    n = 59000
    sigma = (pb * (1 - pb) / n) ** 0.5
    # quantisation to 8 bits, the estimators and the finite window cost a little: allow +25 % and 5 sigma
    assert pb * 0.8 - 5 * sigma < got < pb * 1.25 + 5 * sigma, (got, pb)
