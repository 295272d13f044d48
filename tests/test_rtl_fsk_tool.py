"""the rtl_fsk-shaped receiver tool on the GPU library.

N1  `rtl_fsk --code H_256_512_4 -b` writes the record stream (1 rx_status byte + K/8 payload bytes per modem frame)
    that pirip's own tx/frame_repeater.c consumes ([R tx/frame_repeater.c:55-106], [R script/frame_repeater:36-38]).
    The REAL reference frame_repeater, compiled from /root/reference by oracle/Makefile into oracle/_ref/, is put
    behind the tool and must re-emit the bursts with its burst-control bytes 1 / 0 / 2 ([R tx/rpitx_fsk.cpp:433-497]).
N2  `-s 1800000 -a 40000`: the csdr decimator in front of the modem, one process instead of four ([R README.md:109,172]).
N3  `-u host`: one JSON line per second of signal with the keys script/dash.py reads ([R script/dash.py:26-46]).
"""
import json
import os
import socket
import subprocess

import numpy as np
import pytest

import orc
import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "bin")
REPEATER = os.path.join(ROOT, "oracle", "_ref", "frame_repeater")


def build_tools():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"], stdout=subprocess.DEVNULL)


def coded_bursts(Fs, Rs, name, nbursts, frames_per_burst, ebno_db, seed):
    """IQ of `nbursts` bursts as rpitx_fsk sends them: preamble, frames [UW | data + CRC | parity], silence between
    ([R tx/rpitx_fsk.cpp:313-336,369-423]); returns (u8 IQ, list of payload byte strings per burst)."""
    import pirip_b200 as pb
    m = pb.FskLdpcModem(name, M=2, Rs=Rs, Fs=Fs)
    g = synth.rng(seed)
    preamble = np.array([0, 1] * 25, np.uint8)
    segs, silent, sent = [], [], []
    for b in range(nbursts):
        segs.append(g.integers(0, 2, 700, dtype=np.uint8)); silent.append(True)
        segs.append(preamble); silent.append(False)
        burst = []
        for f in range(frames_per_burst):
            payload = orc.payload_data_bits(m.data_bits).copy()
            payload[0:8] = [0, 0, 0, 0, 0, 0, 0, 1]                     # source address 0x01 ([R tx/rpitx_fsk.cpp:382-386])
            payload[8:16] = [(10 * b + f + 1) >> (7 - i) & 1 for i in range(8)]  # sequence number
            payload = orc.insert_crc(payload)
            burst.append(np.packbits(payload).tobytes())
            segs.append(m.frame(payload)); silent.append(False)
        sent.append(burst)
    segs.append(g.integers(0, 2, 1500, dtype=np.uint8)); silent.append(True)
    x = synth.fsk_mod(np.concatenate(segs), 2, Fs, Rs, Rs, Rs, timing_offset=3.3, phase0=0.7)
    Ts, pos = Fs // Rs, 0
    for s, quiet in zip(segs, silent):
        if quiet:
            x[pos:pos + s.size * Ts] = 0
        pos += s.size * Ts
    x = synth.add_awgn(x, ebno_db, Ts, 2, g)
    m.close()
    return synth.to_u8(x, amp=45.0), sent


def test_burst_records_drive_the_reference_frame_repeater(tmp_path):
    if not os.path.exists(REPEATER):
        pytest.skip("oracle/_ref/frame_repeater not built (needs /root/reference at build time)")
    build_tools()
    Fs, Rs, name = 80000, 10000, "H_256_512_4"
    iq, sent = coded_bursts(Fs, Rs, name, nbursts=2, frames_per_burst=3, ebno_db=9.0, seed=314)
    f = tmp_path / "iq.u8"
    iq.tofile(f)
    rx = subprocess.run([os.path.join(BIN, "rtl_fsk"), "-i", str(f), "-s", str(Fs), "-a", str(Fs), "-r", str(Rs), "-m", "2",
                         "--code", name, "-b", "-q", "-"], capture_output=True, check=True)
    rec = np.frombuffer(rx.stdout, np.uint8).reshape(-1, 33)           # status byte + 256/8 payload bytes per modem frame
    assert rec.shape[0] == (iq.size // 2) // 400                       # one record per modem frame (N = 50 symbols * Ts 8)
    got = [r[1:].tobytes() for r in rec if r[0] & 0x4]                 # FREEDV_RX_BITS
    assert got == [p for burst in sent for p in burst]                 # every frame of both bursts, in order, CRC clean
    assert all(not r[1:].any() for r in rec if not r[0] & 0x4)         # zeros when there is no frame
    # pirip's own state machine behind it: rtl_fsk -b | frame_repeater 256 0x2
    rep = subprocess.run([REPEATER, "256", "0x2"], input=rx.stdout, capture_output=True, check=True, timeout=60)
    out = np.frombuffer(rep.stdout, np.uint8).reshape(-1, 33)
    want_ctl, want_data = [], []
    for burst in sent:
        for k, p in enumerate(burst):
            want_ctl.append(1 if k == 0 else 0)
            want_data.append(bytes([0x02]) + p[1:])                    # source address rewritten to ours
        want_ctl.append(2)
        want_data.append(bytes(32))
    assert out[:, 0].tolist() == want_ctl
    assert [r[1:].tobytes() for r in out] == want_data


def test_source_address_filter_and_payload_stream(tmp_path):
    build_tools()
    Fs, Rs, name = 80000, 10000, "H_256_512_4"
    iq, sent = coded_bursts(Fs, Rs, name, nbursts=1, frames_per_burst=3, ebno_db=10.0, seed=27)
    f = tmp_path / "iq.u8"
    iq.tofile(f)
    base = [os.path.join(BIN, "rtl_fsk"), "-i", str(f), "-a", str(Fs), "-r", str(Rs), "-m", "2", "--code", name, "-q"]
    plain = subprocess.run(base + ["-"], capture_output=True, check=True).stdout
    assert plain == b"".join(sent[0])                                  # packed payload bytes only [R README.md:297]
    filtered = subprocess.run(base + ["--filter", "0x1", "-"], capture_output=True, check=True).stdout
    assert filtered == b""                                             # frames from address 0x01 are dropped [R README.md:303]


def test_uncoded_decimating_front_end_matches_the_csdr_pipeline(tmp_path):
    """rtl_fsk -s 1800000 -a 40000 -r 1000 == csdr convert_u8_f | fir_decimate_cc 45 -> fsk_demod (float samples), and the
    100-bit test packets arrive ([R README.md:109,114])."""
    build_tools()
    Fs_in, Fs, Rs = 1800000, 40000, 1000
    bits = synth.test_bits(1200)
    x = synth.fsk_mod(bits, 2, Fs_in, Rs, 1000, 1000, timing_offset=333.0, phase0=0.3)
    x = synth.add_awgn(x, 12.0, Fs_in // Rs, 2, synth.rng(77))
    iq = synth.to_u8(x, amp=80.0)
    f = tmp_path / "iq.u8"
    iq.tofile(f)
    rx = subprocess.run([os.path.join(BIN, "rtl_fsk"), "-i", str(f), "-s", str(Fs_in), "-a", str(Fs), "-r", str(Rs), "-m", "2",
                         "--fsk_lower", "500", "-q", "-"], capture_output=True, check=True)
    got = np.frombuffer(rx.stdout, np.uint8)
    # oracle: csdr stage in floats, then the demodulator on complex float input
    dec, _ = orc.csdr_decimate_stream(iq, 45)
    # no -p on upstream's command line: P is derived from Ts = 40 (halved while even and above 10 -> 10)
    ref = orc.fsk_demod_stream(orc.demod_cfg(Fs, Rs, 2, P=10, limits=(500, Fs // 2), fmt=orc.FMT_F32), dec)
    assert np.array_equal(got, ref["bits"][:got.size]) and got.size >= ref["bits"].size - 50
    put = subprocess.run([os.path.join(BIN, "fsk_put_test_bits"), "-q", "-p", "8", "-"], input=rx.stdout, capture_output=True)
    assert put.returncode == 0, put.stderr.decode()


def test_dashboard_telemetry_has_the_fields_dash_py_reads(tmp_path):
    build_tools()
    Fs, Rs = 240000, 10000
    bits = synth.test_bits(25000)                                       # 2.5 s of signal: two telemetry lines
    x = synth.fsk_mod(bits, 2, Fs, Rs, 10000, 10000, timing_offset=5.0)
    iq = synth.to_u8(synth.add_awgn(x, 10.0, Fs // Rs, 2, synth.rng(3)), amp=70.0)
    f = tmp_path / "iq.u8"
    iq.tofile(f)
    sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    sock.bind(("127.0.0.1", 8001))
    sock.settimeout(30)
    try:
        subprocess.run([os.path.join(BIN, "rtl_fsk"), "-i", str(f), "-a", str(Fs), "-r", str(Rs), "-m", "2", "-p", "24",
                        "--fsk_lower", "500", "--fsk_upper", "25000", "-u", "127.0.0.1", "-q", os.devnull], check=True)
        lines = [sock.recv(65536) for _ in range(2)]
    finally:
        sock.close()
    for raw in lines:
        d = json.loads(raw)
        assert set(d) >= {"SfdB", "f_est_Hz", "fsk_lower_Hz", "fsk_upper_Hz", "SNRest_lin", "norm_rx_timing", "Fs_Hz"}
        assert len(d["SfdB"]) == 256 and d["Fs_Hz"] == Fs and (d["fsk_lower_Hz"], d["fsk_upper_Hz"]) == (500, 25000)
        assert sorted(round(v / 1000) for v in d["f_est_Hz"]) == [10, 20]   # tones at +10 / +20 kHz
        assert len(d["norm_rx_timing"]) >= 190 and all(-0.5 <= t <= 0.5 for t in d["norm_rx_timing"])
        assert d["SNRest_lin"] > 3


def test_frame_repeater_loopback_in_files(tmp_path):
    """`script/frame_repeater start_loopback` with files for the radios ([R script/frame_repeater:47-64]): pirip's transmitter
    command line (tools/rpitx_fsk: IQ-file build, N4) sends two bursts of three test frames from source 0x01; the GPU
    receiver feeds the reference's own frame_repeater; its output goes through the transmitter again (--packed, burst-control
    bytes) and the receiver gets the same frames back from source 0x02 — and drops them when told to filter 0x2."""
    if not os.path.exists(REPEATER):
        pytest.skip("oracle/_ref/frame_repeater not built (needs /root/reference at build time)")
    build_tools()
    code, Rs, Fs = "H_256_512_4", 10000, 80000
    tx = [os.path.join(BIN, "rpitx_fsk")]
    rx = [os.path.join(BIN, "rtl_fsk"), "-a", str(Fs), "-r", str(Rs), "-m", "2", "--code", code, "-q"]
    a, b = tmp_path / "a.u8", tmp_path / "b.u8"
    subprocess.run(tx + ["/dev/zero", "--code", code, "-r", str(Rs), "-s", str(Rs), "--testframes", "3", "--bursts", "2", "--seq",
                         "--source", "0x1", "--fs", str(Fs), "--ebno", "10", "-o", str(a)], check=True, capture_output=True)
    rec = subprocess.run(rx + ["-i", str(a), "--filter", "0x2", "-b", "-"], check=True, capture_output=True).stdout
    first = [r[1:] for r in np.frombuffer(rec, np.uint8).reshape(-1, 33) if r[0] & 0x4]
    assert len(first) == 6 and all(r[0] == 0x01 for r in first) and [int(r[1]) for r in first] == [1, 2, 3, 1, 2, 3]
    rep = subprocess.run([REPEATER, "256", "0x2"], input=rec, check=True, capture_output=True, timeout=60).stdout
    subprocess.run(tx + ["-", "--code", code, "-r", str(Rs), "-s", str(Rs), "--packed", "--fs", str(Fs), "--ebno", "10", "--seed", "5",
                         "-o", str(b)], input=rep, check=True, capture_output=True)
    back = subprocess.run(rx + ["-i", str(b), "-"], check=True, capture_output=True).stdout
    back = np.frombuffer(back, np.uint8).reshape(-1, 32)
    assert back.shape[0] == 6
    for got, sent in zip(back, first):
        assert got[0] == 0x02 and np.array_equal(got[1:30], sent[1:30])     # same frame, new source address (and new CRC)
    assert subprocess.run(rx + ["-i", str(b), "--filter", "0x2", "-"], check=True, capture_output=True).stdout == b""


def test_transmitter_test_modes(tmp_path):
    """[R README.md:191,196]: --testframes counted by the receiver's own --testframes; [R README.md:101]: uncoded test bits."""
    build_tools()
    code = "H_256_512_4"
    a = tmp_path / "a.u8"
    subprocess.run([os.path.join(BIN, "rpitx_fsk"), "/dev/zero", "--code", code, "-r", "10000", "-s", "10000", "--testframes", "4",
                    "--fs", "80000", "--ebno", "9", "-o", str(a)], check=True, capture_output=True)
    r = subprocess.run([os.path.join(BIN, "rtl_fsk"), "-i", str(a), "-a", "80000", "-r", "10000", "-m", "2", "--code", code,
                        "--testframes", os.devnull], check=True, capture_output=True)
    assert b"4 packets" in r.stderr and b"BER 0.0000" in r.stderr, r.stderr
    bits = subprocess.run([os.path.join(BIN, "fsk_get_test_bits"), "-", "3000"], capture_output=True, check=True).stdout
    t = tmp_path / "t.u8"
    subprocess.run([os.path.join(BIN, "rpitx_fsk"), "-", "-r", "10000", "-s", "10000", "--fs", "240000", "--ebno", "11", "-o", str(t)],
                   input=bits, check=True, capture_output=True)
    dem = subprocess.run([os.path.join(BIN, "rtl_fsk"), "-i", str(t), "-a", "240000", "-r", "10000", "-m", "2", "-p", "24", "-q", "-"],
                         check=True, capture_output=True).stdout
    put = subprocess.run([os.path.join(BIN, "fsk_put_test_bits"), "-q", "-p", "25", "-"], input=dem, capture_output=True)
    assert put.returncode == 0, put.stderr.decode()


@pytest.mark.parametrize("args,Fs,Rs,M", [
    # the reference's own command lines, options exactly as written (tuner options are accepted and ignored; the file source
    # replaces the dongle): [R README.md:114], [R README.md:152], [R README.md:172], [R README.md:196], [R README.md:262],
    # [R README.md:286], [R script/frame_repeater:23], [R test/loopback_rtl_fsk.sh:10]
    ("-g 49 -f 144490000 - -n 480000", 240000, 10000, 2),
    ("-w 500E3 -e ff8 -r 1000 -f 144490000 - -u localhost", 240000, 1000, 2),  # modem rate chosen by the tool: 40 kHz
    ("-s 2400000 -a 80000 -w 500E3 -e ff8 -r 10000 -f 144490000 - -u 127.0.0.1", 80000, 10000, 2),
    ("-g 1 -f 144490000 - -a 100000 -r 10000 --code H_256_512_4 -v -u localhost --testframes", 100000, 10000, 2),
    ("-g 49 -f 144490000 - -a 200000 -r 10000 -m 4 --code H_256_512_4 -v -u localhost --testframes --mask 10000 -e 0xfff", 200000, 10000, 4),
    ("-g 30 -f 144490000 - -r 10000 -m 2 -a 180000 --code H_256_512_4 -v -u localhost --testframes -m 4 --mask 10000", 180000, 10000, 4),
    ("-g 49 -f 144490000 - -a 40000 -r 1000 --code H_256_512_4 -L /dev/null", 40000, 1000, 2),
    ("-g 1 -s 240000 -f 144490000 - -n 240000 -u localhost", 240000, 10000, 2),
])
def test_reference_command_lines_run_verbatim(args, Fs, Rs, M, tmp_path):
    """every rtl_fsk command line of the reference's README / scripts must parse and run (it used to exit with usage()
    or 'Ts/P must be an integer'); the sample source is stdin in place of the dongle."""
    build_tools()
    tuner = Fs if "-s" not in args.split() else int(args.split()[args.split().index("-s") + 1])
    nsec = 0.25 if Rs >= 10000 else 1.0
    x = synth.fsk_mod(synth.test_bits(int(Rs * nsec) * (1 if M == 2 else 2)), M, tuner, Rs, Rs, Rs, timing_offset=3.0)
    iq = synth.to_u8(synth.add_awgn(x, 15.0, tuner // Rs, M, synth.rng(5)), amp=60.0)
    r = subprocess.run([os.path.join(BIN, "rtl_fsk")] + args.split(), input=iq.tobytes(), capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    assert b"modem frames" in r.stderr
    if "--code" not in args:
        assert len(r.stdout) > 0
        if "-n" in args.split():
            n = int(args.split()[args.split().index("-n") + 1])
            assert len(r.stdout) <= n // (Fs // Rs)


def test_chunked_ingest_writes_the_same_stream_as_one_frame_per_call(tmp_path):
    """rtl_fsk --chunk K (default 32: K frames per read and per kernel launch) against --chunk 1 (the fsk_demod.c loop), uncoded
    with the decimating front end and coded with -b records."""
    build_tools()
    Fs_in, Fs, Rs = 2400000, 80000, 10000
    bits = synth.test_bits(6000)
    x = synth.fsk_mod(bits, 2, Fs_in, Rs, 10000, 10000, timing_offset=33.0, phase0=0.3, ppm=60.0)
    iq = synth.to_u8(synth.add_awgn(x, 10.0, Fs_in // Rs, 2, synth.rng(78)), amp=70.0)
    f = tmp_path / "iq.u8"
    iq.tofile(f)
    base = [os.path.join(BIN, "rtl_fsk"), "-i", str(f), "-s", str(Fs_in), "-a", str(Fs), "-r", str(Rs), "-m", "2", "-q"]
    one = subprocess.run(base + ["--chunk", "1", "-"], capture_output=True, check=True).stdout
    for k in ("5", "32"):
        many = subprocess.run(base + ["--chunk", k, "-"], capture_output=True, check=True).stdout
        assert many == one and len(one) >= 5500
    # coded, record stream
    import pirip_b200 as pb
    code = "H_256_512_4"
    m = pb.FskLdpcModem(code, M=2, Rs=10000, Fs=80000)
    frames = [m.frame(orc.insert_crc(orc.payload_data_bits(m.data_bits))) for _ in range(4)]
    m.close()
    tx = np.concatenate([np.zeros(300, np.uint8), np.array([0, 1] * 25, np.uint8)] + frames + [np.zeros(400, np.uint8)])
    y = synth.fsk_mod(tx, 2, 80000, 10000, 10000, 10000, timing_offset=2.2)
    u8 = synth.to_u8(synth.add_awgn(y, 9.0, 8, 2, synth.rng(79)), amp=50.0)
    g = tmp_path / "coded.u8"
    u8.tofile(g)
    cbase = [os.path.join(BIN, "rtl_fsk"), "-i", str(g), "-a", "80000", "-r", "10000", "-m", "2", "--code", code, "-q", "-b"]
    rec1 = subprocess.run(cbase + ["--chunk", "1", "-"], capture_output=True, check=True).stdout
    rec32 = subprocess.run(cbase + ["-"], capture_output=True, check=True).stdout
    assert rec1 == rec32 and len(rec1) % 33 == 0
    assert sum(1 for i in range(0, len(rec1), 33) if rec1[i] & 4) == 4
