"""GPU tests of the boundary: codec2's per-stream struct FSK API, input formats, ragged /
empty / short inputs, batch shapes that do not divide the captures-per-CTA."""
import numpy as np
import pytest

import orc
import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_mode")]

C1 = dict(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000))


def ocfg(kw, **over):
    k = dict(kw)
    k.update(over)
    return orc.demod_cfg(k["Fs"], k["Rs"], k["M"], P=k.get("P", 8), limits=k.get("limits"), mask=k.get("mask", False),
                         tone_spacing=k.get("tone_spacing", 400), fmt=k.get("fmt", orc.FMT_U8), soft=k.get("soft", False))


def test_struct_fsk_stream_api_matches_oracle_and_batch():
    """fsk_create_hbr / fsk_nin / fsk_demod fed like the fsk_demod CLI loop ([R README.md:105])."""
    import pirip_b200 as pb
    u8, _ = synth.capture_u8(21, 1500, 2, 240000, 10000, 10000, 10000, ebno_db=8.0)
    ref = orc.fsk_demod_stream(ocfg(C1), u8)
    x = ((u8.astype(np.float32) - 127.0) / 128.0).view(np.complex64)
    st = pb.FskStream(240000, 10000, 2, P=24, limits=(500, 25000))
    assert (st.s.Ts, st.s.N, st.s.Nmem, st.s.Ndft, st.s.Nbits) == (24, 1200, 1248, 256, 50)
    pos, out, timing = 0, [], []
    while pos + st.nin() <= x.size:
        n = st.nin()
        out.append(st.demod(x[pos:pos + n]))
        timing.append(st.s.norm_rx_timing)
        pos += n
    st.close()
    got = np.concatenate(out)
    assert len(out) == ref["nframes"]
    assert np.array_equal(got, ref["bits"])
    assert np.array_equal(np.array(timing, np.float32).view(np.uint32), ref["stats"]["norm_rx_timing"].view(np.uint32))


def test_struct_fsk_soft_decision_api():
    import pirip_b200 as pb
    kw = dict(Fs=240000, Rs=10000, M=4, P=8, limits=(500, 45000))
    u8, _ = synth.capture_u8(22, 1600, 4, 240000, 10000, 10000, 10000, ebno_db=9.0)
    ref = orc.fsk_demod_stream(ocfg(kw), u8)
    x = ((u8.astype(np.float32) - 127.0) / 128.0).view(np.complex64)
    st = pb.FskStream(240000, 10000, 4, P=8, limits=(500, 45000))
    pos, filt = 0, []
    while pos + st.nin() <= x.size:
        n = st.nin()
        filt.append(st.demod_sd(x[pos:pos + n]))
        pos += n
    st.close()
    assert np.array_equal(np.concatenate(filt).view(np.uint32), ref["rx_filt"].view(np.uint32))


@pytest.mark.parametrize("fmt", ["s16", "f32"])
def test_other_input_formats(fmt):
    import pirip_b200 as pb
    x = synth.fsk_mod(synth.test_bits(1200), 2, 40000, 1000, 1000, 1000, timing_offset=7.3, phase0=1.0)
    x = synth.add_awgn(x, 9.0, 40, 2, synth.rng(77))
    kw = dict(Fs=40000, Rs=1000, M=2, P=8, limits=(500, 20000))
    if fmt == "s16":
        raw, ofmt, pfmt = synth.to_s16(x), orc.FMT_S16, pb.FMT_S16
    else:
        raw, ofmt, pfmt = x.astype(np.complex64), orc.FMT_F32, pb.FMT_F32
    ref = orc.fsk_demod_stream(ocfg(kw, fmt=ofmt), raw)
    plan = pb.FskPlan(**kw, fmt=pfmt)
    out = plan.demod_batch(raw[None, :], want_stats=True)
    nf = ref["nframes"]
    assert out["nframes"][0] == nf and nf > 10
    assert np.array_equal(out["bits"][0, :nf * plan.nbits], ref["bits"])
    assert np.array_equal(out["stats"][0, :nf]["SNRest"].view(np.uint32), ref["stats"]["SNRest"].view(np.uint32))


@pytest.mark.parametrize("ncap", [1, 2, 3, 5, 9])
def test_batch_sizes_not_multiple_of_cta_width(ncap):
    import pirip_b200 as pb
    caps = [synth.capture_u8(500 + i, 700, 2, 240000, 10000, 10000, 10000, ebno_db=[None, 6.0, 10.0][i % 3])[0]
            for i in range(ncap)]
    n = min(len(c) for c in caps)
    iq = np.stack([c[:n] for c in caps])
    plan = pb.FskPlan(**C1)
    out = plan.demod_batch(iq)
    for c in range(ncap):
        ref = orc.fsk_demod_stream(ocfg(C1), iq[c])
        assert out["nframes"][c] == ref["nframes"]
        assert np.array_equal(out["bits"][c, :ref["nframes"] * 50], ref["bits"])


def test_short_odd_and_unaligned_captures():
    """Captures shorter than one frame, one sample short of a frame, odd lengths (row stride not a
    multiple of 16 bytes: the TMA staging falls back to direct loads where it must)."""
    import pirip_b200 as pb
    plan = pb.FskPlan(**C1)
    full, _ = synth.capture_u8(31, 900, 2, 240000, 10000, 10000, 10000, ebno_db=7.0)
    for nsamp in (0, 1, 1199, 1200, 1201, 2407, 3611, 9001):
        if nsamp == 0:
            continue  # zero-length captures are rejected by numpy reshape; covered by n_captures == 0 below
        iq = np.stack([full[:2 * nsamp], full[2:2 * nsamp + 2]])
        out = plan.demod_batch(iq)
        for c in range(2):
            ref = orc.fsk_demod_stream(ocfg(C1), iq[c])
            assert out["nframes"][c] == ref["nframes"], nsamp
            assert np.array_equal(out["bits"][c, :ref["nframes"] * 50], ref["bits"]), nsamp
    out = plan.demod_batch(np.zeros((0, 2400), np.uint8))
    assert out["nframes"].size == 0


def test_all_zero_and_saturated_input_match_oracle():
    """Silence (u8 = 127/128) and full-scale clipping: degenerate statistics (S = N, NaN v_est) must
    come out exactly as the oracle's."""
    import pirip_b200 as pb
    plan = pb.FskPlan(**C1)
    g = synth.rng(9)
    rows = [np.full(2 * 6000, 127, np.uint8), np.full(2 * 6000, 128, np.uint8),
            g.integers(0, 2, 2 * 6000, dtype=np.uint8) * 255, g.integers(0, 256, 2 * 6000, dtype=np.uint8)]
    iq = np.stack(rows)
    out = plan.demod_batch(iq, want_stats=True)
    for c in range(iq.shape[0]):
        ref = orc.fsk_demod_stream(ocfg(C1), iq[c])
        nf = ref["nframes"]
        assert out["nframes"][c] == nf
        assert np.array_equal(out["bits"][c, :nf * 50], ref["bits"]), c
        for name in ("norm_rx_timing", "SNRest", "v_est", "EbNodB", "ppm"):
            assert np.array_equal(out["stats"][c, :nf][name].view(np.uint32), ref["stats"][name].view(np.uint32)), (c, name)


def test_device_resident_entry_point_with_torch_tensors():
    import ctypes

    import torch

    import pirip_b200 as pb
    caps = [synth.capture_u8(60 + i, 900, 2, 240000, 10000, 10000, 10000, ebno_db=8.0)[0] for i in range(7)]
    n = min(len(c) for c in caps) // 16 * 16
    iq = np.stack([c[:n] for c in caps])
    plan = pb.FskPlan(**C1)
    d_iq = torch.from_numpy(iq).cuda()
    mf = plan.max_frames(n // 2)
    d_bits = torch.zeros((7, mf * 50), dtype=torch.uint8, device="cuda")
    d_nf = torch.zeros(7, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    plan.demod_batch_dev(d_iq.data_ptr(), 7, n // 2, d_iq.stride(0), d_bits.data_ptr(), d_nf.data_ptr(), mf,
                         stream=ctypes.c_void_p(s.cuda_stream))
    torch.cuda.synchronize()
    for c in range(7):
        ref = orc.fsk_demod_stream(ocfg(C1), iq[c])
        assert int(d_nf[c]) == ref["nframes"]
        assert np.array_equal(d_bits[c, :ref["nframes"] * 50].cpu().numpy(), ref["bits"])


def test_front_team_pipelined_schedule_is_bit_identical(monkeypatch):
    """PIRIP_FSK_PIPE=1: the frequency estimate runs one frame ahead on a front team of warps."""
    import pirip_b200 as pb
    monkeypatch.setenv("PIRIP_FSK_PIPE", "1")
    caps = [synth.capture_u8(800 + i, 2500, 2, 240000, 10000, 10000, 10000, ebno_db=[None, 5.0, 9.0][i % 3])[0]
            for i in range(5)]
    n = min(len(c) for c in caps)
    iq = np.stack([c[:n] for c in caps])
    plan = pb.FskPlan(**C1)
    out = plan.demod_batch(iq, want_stats=True)
    for c in range(iq.shape[0]):
        ref = orc.fsk_demod_stream(ocfg(C1), iq[c])
        nf = ref["nframes"]
        assert out["nframes"][c] == nf
        assert np.array_equal(out["bits"][c, :nf * 50], ref["bits"])
        assert np.array_equal(out["stats"][c, :nf].view(np.uint8), ref["stats"].view(np.uint8))
    plan.close()


def _both_streams(kw, **extra):
    import pirip_b200 as pb
    return pb.FskStream(**kw, **extra), orc.FskStream(**kw, **extra)


def _run_streams(g, o, x, hook=None, nframes=10 ** 9):
    """Feed both objects the same samples, call for call; every hard bit, nin and norm_rx_timing must agree (uint32 patterns)."""
    pos, k = 0, 0
    while k < nframes and pos + g.nin() <= x.size:
        assert g.nin() == o.nin(), k
        n = g.nin()
        if hook:
            hook(k, g, o)
        bg, bo = g.demod(x[pos:pos + n]), o.demod(x[pos:pos + n])
        assert np.array_equal(bg, bo), k
        assert np.float32(g.s.norm_rx_timing).view(np.uint32) == np.float32(o.s.norm_rx_timing).view(np.uint32), k
        assert [np.float32(v).view(np.uint32) for v in g.s.f_est[:g.s.mode]] == [np.float32(v).view(np.uint32) for v in o.s.f_est[:o.s.mode]], k
        pos += n
        k += 1
    return k


def _signal(seed, nbits, kw, ebno=9.0, ppm=80.0):
    x = synth.fsk_mod(synth.test_bits(nbits), kw["M"], kw["Fs"], kw["Rs"], kw["Rs"], kw["Rs"], timing_offset=2.3, phase0=0.7, ppm=ppm)
    x = synth.add_awgn(x, ebno, kw["Fs"] // kw["Rs"], kw["M"], synth.rng(seed))
    return (0.4 * x).astype(np.complex64)


def test_burst_mode_fixes_nin_and_matches_the_oracle():
    """fsk_enable_burst_mode ([U fsk.c]): nin stays N whatever the timing estimate says (exported but untested)."""
    kw = dict(Fs=80000, Rs=10000, M=2, P=8)
    g, o = _both_streams(kw, limits=(500, 39000))
    g.enable_burst_mode()
    o.enable_burst_mode()
    x = _signal(1, 3000, kw, ppm=400.0)  # a clock offset that would force nin slips within a few frames
    seen = []
    n = _run_streams(g, o, x, hook=lambda k, a, b: seen.append(a.nin()))
    assert n >= 50 and set(seen) == {g.s.N}
    g.close()
    o.close()


def test_lock_nin_and_clear_estimators_match_the_oracle():
    kw = dict(Fs=80000, Rs=10000, M=2, P=8)
    g, o = _both_streams(kw, limits=(500, 39000))
    x = _signal(2, 4000, kw, ppm=400.0)

    def hook(k, a, b):
        if k == 20:  # struct FSK's public field, as rtl_fsk-style callers set it
            a.s.lock_nin = 1
            b.s.lock_nin = 1
        if k == 45:
            a.s.lock_nin = 0
            b.s.lock_nin = 0
        if k == 30:  # fsk_clear_estimators between bursts ([U fsk.c]): Sf and the timing history start again
            a.clear_estimators()
            b.clear_estimators()
    nins = []
    n = _run_streams(g, o, x, hook=lambda k, a, b: (hook(k, a, b), nins.append(a.nin())))
    assert n >= 70
    assert set(nins[21:45]) == {g.s.N} and len(set(nins)) > 1  # locked in between, slipping outside
    g.close()
    o.close()


@pytest.mark.parametrize("Nsym", [32, 100])
def test_other_nsym_stream_and_batch(Nsym):
    """Nsym != 50 through fsk_create_hbr and through the batch plan."""
    import pirip_b200 as pb
    kw = dict(Fs=80000, Rs=10000, M=2, P=8)
    g, o = _both_streams(kw, limits=(500, 39000), Nsym=Nsym)
    assert (g.s.N, g.s.Nbits, g.s.Ndft) == (o.s.N, o.s.Nbits, o.s.Ndft)
    x = _signal(3, 40 * Nsym, kw)
    assert _run_streams(g, o, x) >= 30
    g.close()
    o.close()
    u8, _ = synth.capture_u8(31, 40 * Nsym, 2, 80000, 10000, 10000, 10000, ebno_db=8.0)
    plan = pb.FskPlan(**kw, Nsym=Nsym, limits=(500, 39000))
    out = plan.demod_batch(np.stack([u8, u8[::-1].copy()]))
    ref = orc.fsk_demod_stream(orc.demod_cfg(80000, 10000, 2, P=8, Nsym=Nsym, limits=(500, 39000)), u8)
    nf = int(out["nframes"][0])
    assert nf == ref["nframes"] and np.array_equal(out["bits"][0, :nf * plan.nbits], ref["bits"])
    plan.close()


def test_fsk_mod_ext_vco_tone_frequencies():
    import pirip_b200 as pb
    st = pb.FskStream(80000, 10000, 4, P=8, f1_tx=1200, tone_spacing=800)
    bits = np.array([0, 0, 0, 1, 1, 0, 1, 1], np.uint8)
    vco = st.mod_ext_vco(bits)
    assert vco.size == 4 * 8
    assert vco.reshape(4, 8).tolist() == [[1200.0] * 8, [2000.0] * 8, [2800.0] * 8, [3600.0] * 8]
    st.close()


@pytest.mark.parametrize("nplans,ncap", [(2, 7), (3, 2), (1, 5)])
def test_multi_device_entry_shards_and_keeps_capture_order(nplans, ncap):
    """pirip_fsk_demod_batch_multi (one process, several GPUs): plans on every visible device (several on the
    same device when the box has fewer GPUs than plans), ragged shard sizes, more plans than captures; output in capture order
    and equal to the single-plan call and to the oracle."""
    import torch

    import pirip_b200 as pb
    ndev = torch.cuda.device_count()
    caps = [synth.capture_u8(500 + c, 1500, 2, 240000, 10000, 10000, 10000, ebno_db=[None, 9.0, 6.0][c % 3])[0] for c in range(ncap)]
    n = min(x.size for x in caps)
    iq = np.stack([x[:n] for x in caps])
    plans = [pb.FskPlan(**C1, device=r % ndev) for r in range(nplans)]
    out = pb.demod_batch_multi(plans, iq)
    one = plans[0].demod_batch(iq)
    assert np.array_equal(out["nframes"], one["nframes"]) and np.array_equal(out["bits"], one["bits"])
    for c in (0, ncap - 1):
        ref = orc.fsk_demod_stream(ocfg(C1), iq[c])
        nf = int(out["nframes"][c])
        assert nf == ref["nframes"] and np.array_equal(out["bits"][c, :nf * 50], ref["bits"])
    for p in plans:
        p.close()


@pytest.mark.parametrize("chunk", [7, 64])
def test_many_frames_per_launch_equal_the_frame_per_call_loop(chunk):
    """pirip_fsk_demod_frames: chunks of samples, every whole frame of a chunk in one launch, the rest carried
    over - bits, rx_filt and every per-frame float equal the fsk_demod loop's (and so the oracle's), nin slips included."""
    import pirip_b200 as pb
    kw = dict(Fs=240000, Rs=10000, M=2, P=24)
    u8, _ = synth.capture_u8(77, 9000, 2, 240000, 10000, 10000, 10000, ebno_db=9.0)
    ref = orc.fsk_demod_stream(ocfg(C1), u8)
    x = ((u8.astype(np.float32) - 127.0) / 128.0).view(np.complex64)
    st = pb.FskStream(**kw, limits=(500, 25000))
    N = st.s.N
    buf = np.zeros(0, np.complex64)
    pos, bits, filt, timing = 0, [], [], []
    while True:
        want = chunk * N + 40 - buf.size
        if want > 0 and pos < x.size:
            buf = np.concatenate([buf, x[pos:pos + want]])
            pos += want
        if buf.size < st.nin():
            break
        b, f, s, used = st.demod_frames(buf, chunk)
        assert 1 <= len(b) <= chunk and used <= buf.size
        bits.append(b.ravel())
        filt.append(f.ravel())
        timing.append(s["norm_rx_timing"])
        buf = buf[used:]
    st.close()
    got = np.concatenate(bits)
    assert got.size == ref["bits"].size and np.array_equal(got, ref["bits"])
    assert np.array_equal(np.concatenate(filt).view(np.uint32), ref["rx_filt"].view(np.uint32))
    assert np.array_equal(np.concatenate(timing).view(np.uint32), ref["stats"]["norm_rx_timing"].view(np.uint32))
    assert len(set(ref["stats"]["nin_next"].tolist())) > 1  # the stream slips: chunks meet frames of different lengths
