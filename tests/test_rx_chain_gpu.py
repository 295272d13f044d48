"""GPU parity of the batched coded receive chain (BASELINE.json config 4): u8 captures of LDPC-coded
bursts -> pirip_fsk_ldpc_rx_batch -> per modem frame rx_status + packed payload, against the oracle chain
(orc.fsk_demod_stream -> orc.Deframer -> CRC) capture by capture.  Shapes: `rtl_fsk --code H_256_512_4` [R README.md:184,200-212],
burst layout of the transmitter [R tx/rpitx_fsk.cpp:313-336,369-400], record consumer [R tx/frame_repeater.c:55-62]."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import orc
import synth

pytestmark = pytest.mark.gpu

FS, RS, P = 80000, 10000, 8


def coded_capture(seed, code, M, ebno_db, nbursts=2, frames_per_burst=(1, 4), f1=10000, spacing=10000, nsamp=None):
    """One capture: noise, then `nbursts` bursts (preamble + 1..4 FSK_LDPC frames with CRC) separated by silence."""
    g = synth.rng(seed)
    K = code["CodeLength"] - code["NumberParityBits"]
    bps = 1 if M == 2 else 2
    pre = np.tile(np.arange(M, dtype=np.uint8), 50 // M + 1)[:50]  # 50 symbols cycling through the tones
    preamble = pre[:, None] >> np.arange(bps - 1, -1, -1) & 1
    preamble = preamble.astype(np.uint8).ravel()
    segs, silent, sent = [], [], []
    for b in range(nbursts):
        gap = int(g.integers(300, 900)) * bps
        segs.append(g.integers(0, 2, gap, dtype=np.uint8))
        silent.append(True)
        segs.append(preamble)
        silent.append(False)
        for f in range(int(g.integers(frames_per_burst[0], frames_per_burst[1] + 1))):
            payload = orc.payload_data_bits(K).copy()
            payload[8:24] = g.integers(0, 2, 16, dtype=np.uint8)
            payload = orc.insert_crc(payload)
            sent.append(np.packbits(payload))
            segs.append(orc.fsk_ldpc_frame(code, payload))
            silent.append(False)
    segs.append(g.integers(0, 2, 700 * bps, dtype=np.uint8))
    silent.append(True)
    allbits = np.concatenate(segs)
    Ts = FS // RS
    x = synth.fsk_mod(allbits, M, FS, RS, f1 + g.uniform(-300, 300), spacing, timing_offset=g.uniform(0, Ts), phase0=g.uniform(0, 6.28),
                      ppm=g.uniform(-30, 30))
    pos = 0
    for seg, sil in zip(segs, silent):
        n = seg.size // bps * Ts
        if sil:
            x[pos:pos + n] = 0
        pos += n
    x = synth.add_awgn(x, ebno_db, Ts, M, g)
    u8 = synth.to_u8(x, amp=40.0)
    if nsamp is not None:
        if u8.size < 2 * nsamp:
            u8 = np.concatenate([u8, synth.to_u8(synth.add_awgn(np.zeros(nsamp - u8.size // 2, complex), ebno_db, Ts, M, g), amp=40.0)])
        u8 = u8[:2 * nsamp]
    return u8, sent


def oracle_chain(cfg, code, u8, Nbits, libm=False):
    ref = orc.fsk_demod_stream(cfg, u8)
    od = orc.Deframer(code, max_nbits=2 * Nbits)
    st, pl, info = [], [], []
    for k in range(ref["nframes"]):
        before = od.info()[5]
        s, p = od.push(ref["llr"][k * Nbits:(k + 1) * Nbits])
        i = od.info()
        stepped = before + Nbits >= 32 + code["CodeLength"]
        st.append(s)
        pl.append(np.packbits(p) if s & orc.RX_BITS else None)
        info.append((int(i[1]) if stepped else -1, int(i[2]) if stepped else 0, int(i[3]) if stepped else 0, int(i[4]) if stepped else 0))
    od.close()
    return ref["nframes"], st, pl, info


def make_captures(name, M, n, seed0, ebnos, nsamp, **kw):
    import pirip_b200 as pb
    _, v = pb.ldpc_code(name)
    if M == 4:
        kw.setdefault("f1", 5000)  # 5, 15, 25, 35 kHz: four tones Rs apart below the 40 kHz Nyquist rate
    caps, sent = [], []
    for c in range(n):
        u8, s = coded_capture(seed0 + c, v, M, ebnos[c % len(ebnos)], nsamp=nsamp, **kw)
        caps.append(u8)
        sent.append(s)
    return v, np.stack(caps), sent


@pytest.mark.parametrize("name,M,ncap", [("H_256_512_4", 2, 256), ("HRA_112_112", 2, 64), ("H_256_512_4", 4, 96)])
def test_rx_batch_equals_oracle_chain(name, M, ncap, monkeypatch):
    """Exact demodulator + exact sum-product: status sequence, payload bytes and the -v log fields (uw_loc, uw_err, iter, pcc) of
    every modem frame of every capture equal the oracle chain's."""
    monkeypatch.setenv("PIRIP_FSK_MODE", "exact")
    import pirip_b200 as pb
    spacing = 10000
    nsamp = 52000 if name == "H_256_512_4" else 30000
    if M == 4:
        nsamp //= 2
    v, iq, sent = make_captures(name, M, ncap, 500, [9.0, 7.0, 6.0, 5.0], nsamp)
    kw = dict(Fs=FS, Rs=RS, M=M, P=P, limits=(500, 39000))
    plan = pb.RxPlan(name, **kw, demod_mode="exact", ldpc_mode="sum_product")
    out = plan.rx_batch(iq, want_info=True)
    cfg = orc.demod_cfg(FS, RS, M, P=P, limits=(500, 39000), soft=True)
    with ThreadPoolExecutor(16) as ex:
        refs = list(ex.map(lambda c: oracle_chain(cfg, v, iq[c], plan.nbits), range(ncap)))
    delivered = 0
    total_sent = 0
    for c in range(ncap):
        nf, st, pl, info = refs[c]
        assert int(out["nframes"][c]) == nf, c
        assert out["status"][c, :nf].tolist() == st, (c, out["status"][c, :nf].tolist(), st)
        # uw_err is only meaningful (and only logged by rtl_fsk -v) where a decode was attempted (iter > 0)
        got_info = [(r[0], r[1] if r[2] > 0 else 0, r[2], r[3]) for r in out["info"][c, :nf].tolist()]
        assert got_info == [(i[0], i[1] if i[2] > 0 else 0, i[2], i[3]) for i in info], c
        for k in range(nf):
            if st[k] & orc.RX_BITS:
                assert np.array_equal(out["payload"][c, k], pl[k]), (c, k)
                delivered += any(np.array_equal(pl[k], s) for s in sent[c])
            else:
                assert not out["payload"][c, k].any()
        total_sent += len(sent[c])
    print(name, M, "packets sent", total_sent, "delivered", delivered)
    assert delivered > 0.6 * total_sent
    plan.close()


@pytest.mark.parametrize("ldpc_mode", ["sum_product_fast", "min_sum"])
def test_rx_batch_tolerance_modes_deliver_the_same_packets(ldpc_mode):
    """Fast demodulator + tolerance-mode LDPC update: every packet whose CRC passes is a packet the exact chain also delivers or
    one that was sent; the delivered count stays within a few percent of the exact chain's."""
    import pirip_b200 as pb
    name, M, ncap = "H_256_512_4", 2, 128
    v, iq, sent = make_captures(name, M, ncap, 900, [9.0, 7.0, 6.0, 5.0], 52000)
    kw = dict(Fs=FS, Rs=RS, M=M, P=P, limits=(500, 39000))
    p0 = pb.RxPlan(name, **kw, demod_mode="exact", ldpc_mode="sum_product")
    o0 = p0.rx_batch(iq)
    p0.close()
    p1 = pb.RxPlan(name, **kw, demod_mode="fast", ldpc_mode=ldpc_mode)
    o1 = p1.rx_batch(iq)
    p1.close()
    assert np.array_equal(o0["nframes"], o1["nframes"])
    RXB = pb._lib.RX_BITS
    n0 = n1 = 0
    for c in range(ncap):
        nf = int(o0["nframes"][c])
        sent_set = {s.tobytes() for s in sent[c]}
        for k in range(nf):
            if o0["status"][c, k] & RXB:
                n0 += 1
            if o1["status"][c, k] & RXB:
                n1 += 1
                assert o1["payload"][c, k].tobytes() in sent_set, (c, k)
    print(ldpc_mode, "packets: exact chain", n0, "tolerance chain", n1)
    assert n1 >= (0.97 if ldpc_mode == "sum_product_fast" else 0.85) * n0


def test_per_versus_ebno_sanity_point_4fsk():
    """[R README.md:239-251]: 4FSK + H_256_512_4 loses about 10 % of its packets at Eb/N0 = 6 dB (Eb per INFORMATION bit: rate 1/2,
    two bits per symbol; measured there one dB worse than theory, bursts with acquisition at their start).  The GPU chain's
    packet error rate must cross 10 % inside 6 +- 1 dB.  The code table is the repo's STAND-IN for H_256_512_4 (same (512, 256)
    repeat-accumulate shape, not codec2's matrix: DESIGN.md 5), so this is a sanity point, not a parity claim."""
    import pirip_b200 as pb
    name, M, per_pt = "H_256_512_4", 4, 48
    kw = dict(Fs=FS, Rs=RS, M=M, P=P, limits=(500, 39000))
    plan = pb.RxPlan(name, **kw, demod_mode="exact", ldpc_mode="sum_product")
    per = {}
    for info_ebno in (5.0, 7.0):
        # add_awgn's Eb is per CHANNEL bit: 3 dB below the information bit's at rate 1/2
        v, iq, sent = make_captures(name, M, per_pt, 3000, [info_ebno - 3.0], 40000, nbursts=2, frames_per_burst=(4, 4))
        out = plan.rx_batch(iq)
        tot = got = 0
        for c in range(per_pt):
            ss = {s.tobytes() for s in sent[c]}
            tot += len(sent[c])
            nf = int(out["nframes"][c])
            got += sum(1 for k in range(nf) if out["status"][c, k] & pb._lib.RX_BITS and out["payload"][c, k].tobytes() in ss)
        per[info_ebno] = 1.0 - got / tot
        print("4FSK H_256_512_4 (stand-in) Eb/N0 %.1f dB: %d of %d packets, PER %.3f" % (info_ebno, got, tot, per[info_ebno]))
    plan.close()
    assert per[5.0] > 0.10 > per[7.0], per
