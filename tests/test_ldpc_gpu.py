"""GPU parity of the FEC stage: LDPC sum-product decoder (bits, iteration count and parity-check
count identical to the oracle), the stand-alone LLR map, the FSK_LDPC deframer, and config 4
end to end: framed + coded 2FSK bursts -> demod (soft) -> UW sync -> LDPC -> CRC."""
import numpy as np
import pytest

import orc
import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_mode")]
CODES = ["HRA_112_112", "H_256_512_4", "H_128_256_5"]


def noisy_llrs(cw, ebno_db, g, rate=0.5):
    """BPSK-over-AWGN style LLRs (positive = bit 0) — enough to exercise the decoder."""
    sigma2 = 1.0 / (2 * rate * 10 ** (ebno_db / 10))
    y = (1.0 - 2.0 * cw) + g.standard_normal(cw.shape) * np.sqrt(sigma2)
    return (2.0 * y / sigma2).astype(np.float32)


@pytest.mark.parametrize("name", CODES)
def test_decoder_matches_oracle(name):
    import pirip_b200 as pb
    l, v = pb.ldpc_code(name)
    K = v["CodeLength"] - v["NumberParityBits"]
    g = synth.rng(17)
    cws, llrs = [], []
    for i, eb in enumerate([8.0, 4.0, 3.0, 2.0, 1.5, 1.0, 0.5, 0.0, -1.0, -3.0] * 3):
        ib = g.integers(0, 2, K, dtype=np.uint8)
        cw = np.concatenate([ib, pb.ldpc_encode(l, ib)])
        cws.append(cw)
        llrs.append(noisy_llrs(cw, eb, g))
    llrs = np.stack(llrs)
    llrs[5, ::7] = 0.0       # exact zeros (erasures)
    llrs[6] *= 1e-6          # tiny magnitudes: the x < 9.08e-5 branch of phi0
    llrs[7] *= 50.0          # large magnitudes: the x > 10 branch
    plan = pb.LdpcPlan(l)
    bits, it, pcc = plan.decode(llrs)
    n_conv = 0
    for i in range(llrs.shape[0]):
        rb, rit, rpcc = orc.ldpc_decode(v, llrs[i])
        assert np.array_equal(bits[i], rb), (name, i)
        assert (it[i], pcc[i]) == (rit, rpcc), (name, i, it[i], pcc[i], rit, rpcc)
        n_conv += int(rpcc == v["NumberParityBits"])
    assert 5 < n_conv < llrs.shape[0]  # the set covers both converged and failed decodes
    plan.close()


@pytest.mark.parametrize("mode", ["sum_product_fast", "min_sum"])
@pytest.mark.parametrize("name", ["H_256_512_4", "HRA_112_112"])
def test_tolerance_decoders_same_payload_when_converged(name, mode):
    """The single-precision sum-product and the normalised min-sum update: on every word BOTH decoders bring to a valid
    codeword the payload bits are the oracle's; the iteration cap ([R README.md:211]: 15) is honoured; the word error rate
    over an Eb/N0 sweep stays within a small factor of the exact decoder's (min-sum breaks down a few tenths of a dB earlier)."""
    import pirip_b200 as pb
    l, v = pb.ldpc_code(name)
    N, NPB = v["CodeLength"], v["NumberParityBits"]
    K = N - NPB
    g = synth.rng(23)
    ebnos = [4.0, 3.0, 2.5, 2.0, 1.5, 1.0]
    per_pt = 300
    ibs = g.integers(0, 2, (len(ebnos) * per_pt, K), dtype=np.uint8)
    cws = np.stack([np.concatenate([ib, pb.ldpc_encode(l, ib)]) for ib in ibs])
    llrs = np.stack([noisy_llrs(cws[i], ebnos[i // per_pt], g) for i in range(cws.shape[0])])
    plan = pb.LdpcPlan(l)
    b0, it0, pcc0 = plan.decode(llrs)
    plan.set_mode(mode)
    b1, it1, pcc1 = plan.decode(llrs)
    plan.close()
    assert it1.max() <= l.max_iter and it1.min() >= 1
    both = (pcc0 == NPB) & (pcc1 == NPB)
    assert both.sum() > (0.3 if mode == "min_sum" else 0.5) * len(both)  # min-sum gives up earlier on the short code
    # two decoders that both end on a codeword can in principle end on different codewords; at these SNRs they do not
    assert np.array_equal(b0[both][:, :K], b1[both][:, :K])
    # converged words carry the transmitted payload, except for the code's own undetected errors (a property of the stand-in
    # tables at 1 - 2 dB, the same words in the exact decoder): rare, and no more of them than the exact decoder has
    und0 = int((b0[pcc0 == NPB][:, :K] != ibs[pcc0 == NPB]).any(axis=1).sum())
    und1 = int((b1[pcc1 == NPB][:, :K] != ibs[pcc1 == NPB]).any(axis=1).sum())
    print(name, mode, "undetected word errors: exact", und0, mode, und1, "of", len(ibs))
    assert und1 <= und0 + 3 and und1 <= 0.01 * len(ibs)
    for k, eb in enumerate(ebnos):
        sl = slice(k * per_pt, (k + 1) * per_pt)
        per0 = float((pcc0[sl] != NPB).mean())
        per1 = float((pcc1[sl] != NPB).mean())
        print(name, mode, "Eb/N0", eb, "PER exact", per0, mode, per1)
        assert per1 <= max(3.0 * per0 + 0.03, 0.02) if mode == "min_sum" else per1 <= 1.3 * per0 + 0.02, (eb, per0, per1)


def test_run_ldpc_decoder_codec2_entry_point():
    import ctypes as C

    import pirip_b200 as pb
    L = pb._lib._ldpc_protos()
    l, v = pb.ldpc_code("H_256_512_4")
    g = synth.rng(3)
    ib = g.integers(0, 2, 256, dtype=np.uint8)
    cw = np.concatenate([ib, pb.ldpc_encode(l, ib)])
    llr = noisy_llrs(cw, 2.5, g)
    out = np.zeros(512, np.uint8)
    pcc = C.c_int(0)
    it = L.run_ldpc_decoder(C.byref(l), out.ctypes.data, llr.ctypes.data, C.byref(pcc))
    rb, rit, rpcc = orc.ldpc_decode(v, llr)
    assert np.array_equal(out, rb) and (it, pcc.value) == (rit, rpcc)


@pytest.mark.parametrize("M", [2, 4])
def test_fsk_rx_filt_to_llrs_entry_point(M):
    import pirip_b200 as pb
    L = pb._lib._ldpc_protos()
    g = synth.rng(M)
    nsym = 50
    rx = np.abs(g.standard_normal(M * nsym)).astype(np.float32) * 3
    for v_est, snr in ((1.7, 6.0), (0.3, 0.9), (2.0, 400.0)):
        out = np.zeros(nsym * (1 if M == 2 else 2), np.float32)
        L.fsk_rx_filt_to_llrs(out.ctypes.data, rx.ctypes.data, np.float32(v_est), np.float32(snr), M, nsym)
        ref = orc.rx_filt_to_llrs(rx, v_est, snr, M, nsym)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def burst_bits(modem, v, nframes, seed, seq=True):
    frames = []
    for f in range(nframes):
        payload = orc.payload_data_bits(modem.data_bits).copy()
        if seq:
            payload[8:16] = [(f + 1 + seed) >> (7 - i) & 1 for i in range(8)]
        frames.append(modem.frame(orc.insert_crc(payload)))
    return frames


@pytest.mark.parametrize("name", CODES[:2])
def test_deframer_matches_oracle_on_llr_stream(name):
    import pirip_b200 as pb
    m = pb.FskLdpcModem(name)
    _, v = pb.ldpc_code(name)
    g = synth.rng(23)
    frames = burst_bits(m, v, 5, 0)
    stream = np.concatenate([g.integers(0, 2, 137, dtype=np.uint8)] + frames + [g.integers(0, 2, 700, dtype=np.uint8)])
    llr = noisy_llrs(stream, 3.0, g)
    od = orc.Deframer(v, max_nbits=100)
    got, want = [], []
    for k in range(0, llr.size - 50, 50):
        st, pl = m.rx(llr[k:k + 50])
        ost, opl = od.push(llr[k:k + 50])
        got.append(st)
        want.append(ost)
        if ost & orc.RX_BITS:
            assert np.array_equal(pl, opl)
    assert got == want
    assert sum(1 for s in got if s & pb._lib.RX_BITS) >= 4
    od.close()
    m.close()


def test_config4_end_to_end_coded_bursts():
    """C4: 2FSK Rs=10k at modem Fs=80k (rtl_fsk -a 80000 -r 10000, [R README.md:172]), bursts of 3 frames
    of H_256_512_4 with preamble and inter-burst noise; GPU soft demod -> deframer -> LDPC -> CRC must
    equal the oracle chain frame for frame, and deliver the packets."""
    import pirip_b200 as pb
    Fs, Rs, M, P = 80000, 10000, 2, 8
    name = "H_256_512_4"
    m = pb.FskLdpcModem(name, M=M, Rs=Rs, Fs=Fs)
    _, v = pb.ldpc_code(name)
    g = synth.rng(99)
    preamble = np.array([0, 1] * 25, np.uint8)  # 50*(M/2) symbols cycling 0,1 ([R tx/rpitx_fsk.cpp:313-336])
    bits = []
    for b in range(2):
        bits.append(g.integers(0, 2, 600, dtype=np.uint8))  # no carrier: replaced by noise below
        bits.append(preamble)
        bits += burst_bits(m, v, 3, 10 * b)
    bits.append(g.integers(0, 2, 900, dtype=np.uint8))
    allbits = np.concatenate(bits)
    x = synth.fsk_mod(allbits, M, Fs, Rs, 10000, 10000, timing_offset=2.7, phase0=0.2)
    Ts = Fs // Rs
    # silence between bursts: zero the carrier where the filler bits are
    pos = 0
    for seg in bits:
        n = seg.size * Ts
        if seg.size in (600, 900):
            x[pos:pos + n] = 0
        pos += n
    x = synth.add_awgn(x, 7.0, Ts, M, g)
    u8 = synth.to_u8(x, amp=45.0)
    kw = dict(Fs=Fs, Rs=Rs, M=M, P=P, limits=(500, 35000), soft=True)
    plan = pb.FskPlan(**kw)
    out = plan.demod_batch(u8[None, :], want_llr=True)
    nf = int(out["nframes"][0])
    cfg = orc.demod_cfg(Fs, Rs, M, P=P, limits=(500, 35000), soft=True)
    ref = orc.fsk_demod_stream(cfg, u8)
    assert nf == ref["nframes"]
    assert np.array_equal(out["llr"][0, :nf * 50].view(np.uint32), ref["llr"].view(np.uint32))
    od = orc.Deframer(v, max_nbits=100)
    npk = 0
    for k in range(nf):
        st, pl = m.rx(out["llr"][0, k * 50:(k + 1) * 50])
        ost, opl = od.push(ref["llr"][k * 50:(k + 1) * 50])
        assert st == ost, (k, st, ost)
        if st & pb._lib.RX_BITS:
            assert np.array_equal(pl, opl)
            npk += 1
    assert npk >= 5, npk  # 6 packets sent at Eb/N0 = 7 dB (rate 1/2): nearly all must arrive
    od.close()
    m.close()
