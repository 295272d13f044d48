"""The command-line tools in the shape of pirip's own loopback tests: a file stands in for rtl_sdr,
everything else is the pipeline of [R test/loopback_rtl_sdr.sh:16] / [R README.md:105,109] and the
pass criterion of [R test/loopback_rtl_sdr.sh:7-10] (>= 90 of 100 100-bit packets)."""
import os
import subprocess

import numpy as np
import pytest

import orc
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "bin")


def build_tools():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"], stdout=subprocess.DEVNULL)


def test_get_put_test_bits_without_gpu(tmp_path):
    build_tools()
    bits = subprocess.run([os.path.join(BIN, "fsk_get_test_bits"), "-", "1000"], capture_output=True, check=True).stdout
    arr = np.frombuffer(bits, np.uint8)
    assert arr.size == 1000 and np.array_equal(arr[:100], synth.TEST_FRAME_100) and np.array_equal(arr[100:200], arr[:100])
    ok = subprocess.run([os.path.join(BIN, "fsk_put_test_bits"), "-q", "-p", "9", "-"], input=bits[37:], capture_output=True)
    assert ok.returncode == 0 and b"PASS" in ok.stderr  # 9 whole packets survive a 37-bit offset
    corrupted = bytearray(bits)
    for i in range(0, 1000, 3):
        corrupted[i] ^= 1
    bad = subprocess.run([os.path.join(BIN, "fsk_put_test_bits"), "-q", "-p", "5", "-"], input=bytes(corrupted), capture_output=True)
    assert bad.returncode == 1 and b"FAIL" in bad.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [False, True])
def test_loopback_rtl_sdr_shaped_pipeline(tmp_path, batch):
    """fsk_get_test_bits | <modulator + channel, file for rtl_sdr> | fsk_demod -d -p 24 2 Fs Rb | fsk_put_test_bits -q -p 90"""
    build_tools()
    Fs, Rb = 240000, 10000  # [R test/include.sh:1-2]
    bits = np.frombuffer(subprocess.run([os.path.join(BIN, "fsk_get_test_bits"), "-", "10000"], capture_output=True,
                                        check=True).stdout, np.uint8)
    x = synth.fsk_mod(bits, 2, Fs, Rb, 2 * Rb, 2 * Rb, timing_offset=9.4, phase0=1.1)  # rx tuned 2*Rb low, shift 2*Rs
    x = synth.add_awgn(x, 11.0, Fs // Rb, 2, synth.rng(5))
    iq = synth.to_u8(x, amp=70.0)
    f = tmp_path / "iq.u8"
    iq.tofile(f)
    cmd = [os.path.join(BIN, "fsk_demod"), "-d", "-p", "24"] + (["--batch"] if batch else []) + ["2", str(Fs), str(Rb), str(f), "-"]
    dem = subprocess.run(cmd, capture_output=True, check=True)
    rx = np.frombuffer(dem.stdout, np.uint8)
    ref = orc.fsk_demod_stream(orc.demod_cfg(Fs, Rb, 2, P=24, limits=(-Fs // 2, Fs // 2)), iq)  # complex input, no --fsk_lower: whole spectrum
    assert np.array_equal(rx, ref["bits"])  # the CLI output is the oracle's, byte for byte
    put = subprocess.run([os.path.join(BIN, "fsk_put_test_bits"), "-q", "-p", "90", "-"], input=dem.stdout, capture_output=True)
    assert put.returncode == 0, put.stderr.decode()


@pytest.mark.gpu
def test_csdr_pipeline_matches_oracle(tmp_path):
    """csdr convert_u8_f | csdr fir_decimate_cc 45 | csdr convert_f_s16 | fsk_demod --fsk_lower 500 -c 2 40000 1000"""
    build_tools()
    bits = synth.test_bits(300)
    x = synth.fsk_mod(bits, 2, 1800000, 1000, 1000, 1000, phase0=0.4)
    iq = synth.to_u8(x, amp=60.0)
    p1 = subprocess.run([os.path.join(BIN, "csdr"), "convert_u8_f"], input=iq.tobytes(), capture_output=True, check=True)
    p2 = subprocess.run([os.path.join(BIN, "csdr"), "fir_decimate_cc", "45"], input=p1.stdout, capture_output=True, check=True)
    p3 = subprocess.run([os.path.join(BIN, "csdr"), "convert_f_s16"], input=p2.stdout, capture_output=True, check=True)
    s16 = np.frombuffer(p3.stdout, np.int16)
    _, ref_s16 = orc.csdr_decimate_stream(iq, 45)
    assert np.array_equal(s16, ref_s16)
    f = tmp_path / "d.s16"
    s16.tofile(f)
    dem = subprocess.run([os.path.join(BIN, "fsk_demod"), "--fsk_lower", "500", "-c", "2", "40000", "1000", str(f), "-"],
                         capture_output=True, check=True)
    ref = orc.fsk_demod_stream(orc.demod_cfg(40000, 1000, 2, P=8, limits=(500, 20000), fmt=orc.FMT_S16), ref_s16)
    assert np.array_equal(np.frombuffer(dem.stdout, np.uint8), ref["bits"])
    ber, n = synth.ber_vs_stream(np.frombuffer(dem.stdout, np.uint8))
    assert n >= 100 and ber == 0.0


@pytest.mark.gpu
def test_fsk_demod_stats_lines(tmp_path):
    """fsk_demod -t / --stats: one JSON object per N frames on stderr with upstream's keys; the bit stream is unchanged"""
    import json
    build_tools()
    Fs, Rb = 240000, 10000
    bits = synth.test_bits(3000)
    x = synth.add_awgn(synth.fsk_mod(bits, 2, Fs, Rb, 10000, 10000, timing_offset=3.0), 12.0, Fs // Rb, 2, synth.rng(8))
    iq = synth.to_u8(x, amp=70.0)
    f = tmp_path / "iq.u8"
    iq.tofile(f)
    base = [os.path.join(BIN, "fsk_demod"), "-d", "-p", "24", "--fsk_lower", "500", "--fsk_upper", "25000"]
    plain = subprocess.run(base + ["2", str(Fs), str(Rb), str(f), "-"], capture_output=True, check=True)
    st = subprocess.run(base + ["-t4", "2", str(Fs), str(Rb), str(f), "-"], capture_output=True, check=True)
    assert st.stdout == plain.stdout and plain.stderr == b""
    lines = [json.loads(ln) for ln in st.stderr.decode().splitlines()]
    nframes = len(plain.stdout) // 50
    assert len(lines) == nframes // 4 and nframes >= 40
    ref = orc.fsk_demod_stream(orc.demod_cfg(Fs, Rb, 2, P=24, limits=(500, 25000)), iq)
    for k, ln in enumerate(lines):
        assert set(ln) >= {"secs", "EbNodB", "ppm", "f1_est", "f2_est", "eye_diagram", "samp_fft"}
        fr = 4 * (k + 1) - 1  # the frame the line was printed after
        assert abs(ln["f1_est"] - float(ref["stats"]["f_est"][fr][0])) < 0.06 and abs(ln["f2_est"] - float(ref["stats"]["f_est"][fr][1])) < 0.06
    assert abs(lines[-1]["f1_est"] - 10000) < 1000 and abs(lines[-1]["f2_est"] - 20000) < 1000
    st8 = subprocess.run(base + ["--stats", "2", str(Fs), str(Rb), str(f), "-"], capture_output=True, check=True)
    assert len(st8.stderr.decode().splitlines()) == nframes // 8
