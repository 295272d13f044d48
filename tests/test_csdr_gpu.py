"""GPU parity of the fused csdr stage (convert_u8_f | fir_decimate_cc | convert_f_s16) and of the
whole config-2 chain: u8 IQ at 1.8 MHz -> /45 -> fsk_demod -c 2 40000 1000 ([R README.md:109])."""
import numpy as np
import pytest

import orc
import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_mode")]


def test_taps_match_oracle_bit_for_bit():
    import pirip_b200 as pb
    for D, tbw in ((45, 0.05), (50, 0.05), (30, 0.05), (8, 0.1)):
        p = pb.DecimPlan(D, tbw)
        ref = orc.lib()
        n = ref.orc_firdes_filter_len(tbw)
        taps = np.zeros(n, np.float32)
        ref.orc_firdes_lowpass_f(orc._ptr(taps), n, np.float32(0.5 / np.float32(D)))
        assert p.taps.size == n
        assert np.array_equal(p.taps.view(np.uint32), taps.view(np.uint32))
        p.close()


@pytest.mark.parametrize("D,n", [(45, 45 * 300 + 79), (45, 100003), (30, 50001), (45, 78), (45, 79), (7, 4099)])
def test_decimator_matches_oracle(D, n):
    import pirip_b200 as pb
    g = synth.rng(D * 1000 + n % 997)
    iq = g.integers(0, 256, (3, 2 * n), dtype=np.uint8)
    p = pb.DecimPlan(D)
    f, s = p.run(iq)
    for r in range(3):
        yf, ys = orc.csdr_decimate_stream(iq[r], D)
        assert f.shape[1] == yf.size
        assert np.array_equal(f[r].view(np.uint32), yf.view(np.uint32))
        assert np.array_equal(s[r], ys)
    p.close()


def test_libcsdr_named_fir_decimate_cc():
    import ctypes as C

    import pirip_b200 as pb
    L = pb.lib()
    g = synth.rng(4)
    x = (g.standard_normal(5000) + 1j * g.standard_normal(5000)).astype(np.complex64)
    taps = orc.firdes_lowpass(45)
    out = np.zeros(200, np.complex64)
    L.fir_decimate_cc.restype = C.c_int
    L.fir_decimate_cc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = L.fir_decimate_cc(x.ctypes.data, out.ctypes.data, x.size, 45, taps.ctypes.data, taps.size)
    ref = np.zeros(200, np.complex64)
    nref = orc.lib().orc_fir_decimate_cc(orc._ptr(x), orc._ptr(ref), x.size, 45, orc._ptr(taps), taps.size)
    assert n == nref
    assert np.array_equal(out[:n].view(np.uint32), ref[:n].view(np.uint32))


def test_config2_chain_decimate_then_demod():
    """C2: 2-FSK at Rs=1000 sampled at 1.8 MHz, csdr /45, fsk_demod --fsk_lower 500 -c 2 40000 1000."""
    import pirip_b200 as pb
    Fs_in, D, Fs, Rs = 1800000, 45, 40000, 1000
    caps = []
    for i, eb in enumerate([None, 10.0]):
        bits = synth.test_bits(400)
        x = synth.fsk_mod(bits, 2, Fs_in, Rs, 1000, 1000, timing_offset=100.0 * i, phase0=0.5)
        x = synth.add_awgn(x, eb, Fs_in // Rs, 2, synth.rng(70 + i))
        caps.append(synth.to_u8(x, amp=60.0))
    nmin = min(len(c) for c in caps)
    iq = np.stack([c[:nmin] for c in caps])
    dec = pb.DecimPlan(D)
    _, s16 = dec.run(iq, want_f32=False)
    plan = pb.FskPlan(Fs, Rs, 2, P=8, limits=(500, 20000), fmt=pb.FMT_S16)
    out = plan.demod_batch(s16, want_stats=True)
    cfg = orc.demod_cfg(Fs, Rs, 2, P=8, limits=(500, 20000), fmt=orc.FMT_S16)
    for c in range(2):
        _, ref_s16 = orc.csdr_decimate_stream(iq[c], D)
        ref = orc.fsk_demod_stream(cfg, ref_s16)
        nf = ref["nframes"]
        assert nf >= 7 and out["nframes"][c] == nf
        assert np.array_equal(out["bits"][c, :nf * 50], ref["bits"])
        assert np.array_equal(out["stats"][c, :nf]["norm_rx_timing"].view(np.uint32),
                              ref["stats"]["norm_rx_timing"].view(np.uint32))
        ber, n = synth.ber_vs_stream(out["bits"][c, :nf * 50])
        assert ber < 0.06  # few hundred bits at 10 dB through the aliasing-prone csdr default filter


@pytest.mark.parametrize("D,n,rows", [(45, 45 * 16 * 40 + 79, 5), (45, 45 * 16 * 3 + 78 + 45 * 15, 2), (30, 30 * 15 * 7 + 79 + 30 * 4, 4),
                                      (30, 30 * 15 * 2 + 79, 1), (45, 45 * 15 + 79, 3)])
def test_register_streaming_kernel_vs_generic_kernel_and_oracle(D, n, rows):
    """fir_decimate_u8_fast (D = 45 / 30, 79 taps: whole groups of 16 / 15 outputs per thread, halo blocks, the
    per-stream remainder loop) against the generic staged kernel (PIRIP_DECIM_GENERIC=1) and the oracle."""
    import os

    import pirip_b200 as pb
    g = synth.rng(D * 7 + n % 1013)
    iq = g.integers(0, 256, (rows, 2 * n), dtype=np.uint8)
    p = pb.DecimPlan(D)
    f, s = p.run(iq)
    os.environ["PIRIP_DECIM_GENERIC"] = "1"
    try:
        fg, sg = p.run(iq)
    finally:
        del os.environ["PIRIP_DECIM_GENERIC"]
    assert np.array_equal(f.view(np.uint32), fg.view(np.uint32)) and np.array_equal(s, sg)
    for r in range(rows):
        yf, ys = orc.csdr_decimate_stream(iq[r], D)
        assert np.array_equal(f[r].view(np.uint32), yf.view(np.uint32)) and np.array_equal(s[r], ys)
    # extreme byte values: the table rows at both ends
    iq2 = np.where(g.integers(0, 2, iq.shape, dtype=np.uint8) == 1, 255, 0).astype(np.uint8)
    f2, s2 = p.run(iq2)
    yf, ys = orc.csdr_decimate_stream(iq2[0], D)
    assert np.array_equal(f2[0].view(np.uint32), yf.view(np.uint32)) and np.array_equal(s2[0], ys)
    p.close()
