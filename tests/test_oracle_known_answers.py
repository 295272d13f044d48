"""CPU tests: pin the oracle with every known-answer that does not depend on the (absent)
codec2/csdr sources (SURVEY.md 8c "self-derivable known-answer checks")."""
import numpy as np
import pytest

import orc
import synth


def test_crc16_ccitt_false_check_value():
    # CRC-16/CCITT-FALSE("123456789") = 0x29B1 — freedv_gen_crc16 as used at [R tx/rpitx_fsk.cpp:80]
    assert orc.crc16(b"123456789") == 0x29B1
    assert orc.crc16(b"") == 0xFFFF


def test_uw_is_0x5186fe15_msb_first():
    assert "".join(map(str, orc.uw_bits())) == "01010001100001101111111000010101"


def test_payload_lcg_bits():
    # ofdm_generate_payload_data_bits [R tx/rpitx_fsk.cpp:367], first 64
    assert "".join(map(str, orc.payload_data_bits(64))) == \
        "1010110001010111100110001110100010011100100000010110001010101110"


def test_fsk_get_test_bits_frame_matches_glibc_constant():
    assert np.array_equal(orc.test_bits_frame(100), synth.TEST_FRAME_100)


def test_pack_unpack_crc_insert():
    bits = orc.payload_data_bits(256)
    with_crc = orc.insert_crc(bits)
    assert np.array_equal(with_crc[:240], bits[:240])
    L = orc.lib()
    assert L.orc_check_crc16_unpacked(orc._ptr(with_crc), 256) == 1
    with_crc[7] ^= 1
    assert L.orc_check_crc16_unpacked(orc._ptr(with_crc), 256) == 0


def test_csdr_taps_properties():
    taps = orc.firdes_lowpass(45)
    # C semantics of firdes_filter_len with a float 0.05: (int)(4.0/0.05f) = 79, already odd
    assert taps.size == 79
    assert abs(float(taps.sum()) - 1.0) < 1e-6
    assert np.allclose(taps, taps[::-1], atol=1e-9)
    assert (taps > 0).all()
    # frequency response at the input rate of config 2 (1.8 MHz): unity at DC, ~-4 dB at the new
    # Nyquist (20 kHz), strong attenuation beyond 60 kHz
    Fs = 1.8e6
    H = lambda f: abs(np.sum(taps.astype(np.float64) * np.exp(-2j * np.pi * f / Fs * np.arange(taps.size))))
    assert abs(20 * np.log10(H(1e3))) < 0.05
    assert -4.5 < 20 * np.log10(H(20e3)) < -3.3
    assert 20 * np.log10(H(60e3)) < -40


def test_csdr_decimate_stream_matches_numpy():
    g = synth.rng(5)
    u8 = g.integers(0, 256, 2 * 5000, dtype=np.uint8)
    y, s16 = orc.csdr_decimate_stream(u8, 45)
    x = (u8.astype(np.float64) / 127.5 - 1.0).view(np.float64)
    xc = x[0::2] + 1j * x[1::2]
    taps = orc.firdes_lowpass(45).astype(np.float64)
    nout = (xc.size - taps.size) // 45 + 1
    ref = np.array([np.dot(xc[45 * k:45 * k + taps.size], taps) for k in range(nout)])
    assert y.size == nout
    assert np.allclose(y, ref, atol=2e-6)
    assert np.array_equal(s16[0::2], np.trunc(y.real.astype(np.float32) * np.float32(32767)).astype(np.int16))


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024])
def test_fft_matches_numpy(n):
    g = synth.rng(n)
    x = (g.standard_normal(n) + 1j * g.standard_normal(n)).astype(np.complex64)
    X = orc.fft(x)
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.max(np.abs(X - ref)) < 2e-5 * np.sqrt(n) * np.max(np.abs(ref))


@pytest.mark.parametrize("M,P,f1,sp", [(2, 24, 10000, 10000), (2, 8, 20000, 20000), (4, 8, 10000, 10000)])
def test_noise_free_zero_errors_any_timing_offset(M, P, f1, sp):
    Fs, Rs = 240000, 10000
    cfg = orc.demod_cfg(Fs, Rs, M, P=P, limits=(500, 60000))
    nbits = 4000 * (2 if M == 4 else 1)
    bits = synth.test_bits(nbits)
    for toff in (0.0, 5.5, 11.9, 17.3, 23.0):
        x = synth.fsk_mod(bits, M, Fs, Rs, f1, sp, timing_offset=toff, phase0=0.3)
        r = orc.fsk_demod_stream(cfg, synth.to_u8(x))
        ber, n = synth.ber_vs_stream(r["bits"])
        assert n > 2000 and ber == 0.0, (toff, ber)


def test_clock_offset_exercises_nin_slips():
    Fs, Rs = 240000, 10000
    cfg = orc.demod_cfg(Fs, Rs, 2, P=24, limits=(500, 25000))
    bits = synth.test_bits(20000)
    x = synth.fsk_mod(bits, 2, Fs, Rs, 10000, 10000, ppm=400.0, timing_offset=3.0)
    r = orc.fsk_demod_stream(cfg, synth.to_u8(x))
    nins = set(int(v) for v in r["stats"]["nin_next"])
    assert len(nins) >= 2, nins  # the timing loop had to slip at least once
    ber, n = synth.ber_vs_stream(r["bits"])
    assert ber < 1e-3


@pytest.mark.parametrize("ebno,tol", [(9.0, 0.35), (6.0, 0.25)])
def test_2fsk_ber_tracks_noncoherent_theory(ebno, tol):
    # noncoherent orthogonal 2FSK: Pb = 0.5 exp(-Eb/2N0)
    Fs, Rs = 240000, 10000
    cfg = orc.demod_cfg(Fs, Rs, 2, P=24, limits=(500, 25000))
    errs = tot = 0
    for c in range(4):
        u8, _ = synth.capture_u8(300 + c, 20000, 2, Fs, Rs, 10000, 10000, ebno_db=ebno)
        r = orc.fsk_demod_stream(cfg, u8)
        ber, n = synth.ber_vs_stream(r["bits"])
        errs += ber * n
        tot += n
    theory = 0.5 * np.exp(-(10 ** (ebno / 10)) / 2)
    assert abs(errs / tot - theory) / theory < tol, (errs / tot, theory)


def test_soft_decisions_agree_with_hard_bits():
    Fs, Rs = 240000, 10000
    cfg = orc.demod_cfg(Fs, Rs, 4, P=8, limits=(500, 45000), soft=True)
    u8, _ = synth.capture_u8(11, 6000, 4, Fs, Rs, 10000, 10000, ebno_db=9.0)
    r = orc.fsk_demod_stream(cfg, u8)
    hard_from_llr = (r["llr"] < 0).astype(np.uint8)  # llr > 0 means bit 0
    agree = (hard_from_llr == r["bits"]).mean()
    assert agree > 0.999, agree


def test_short_and_empty_inputs():
    cfg = orc.demod_cfg(240000, 10000, 2, P=24)
    assert orc.fsk_demod_stream(cfg, np.zeros(0, np.uint8))["nframes"] == 0
    assert orc.fsk_demod_stream(cfg, np.full(2 * 1199, 127, np.uint8))["nframes"] == 0  # one sample short
    assert orc.fsk_demod_stream(cfg, np.full(2 * 1200, 127, np.uint8))["nframes"] == 1
