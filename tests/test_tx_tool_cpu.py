"""CPU-only: the IQ-file build of pirip's transmitter (tools/rpitx_fsk) against the ORACLE receive chain
(oracle fsk_demod soft decisions -> oracle deframer / LDPC / CRC).  No GPU: the tool only uses the host-side framer,
CRC and pack entry points of libpirip_b200.so, and the receiver here is the oracle, so the wire format the transmitter
produces ([R tx/rpitx_fsk.cpp:75-83,313-336,369-423]) is pinned independently of the CUDA path."""
import os
import subprocess

import numpy as np

import orc
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "bin")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s", "bin/rpitx_fsk", "bin/fsk_get_test_bits"],
                          stdout=subprocess.DEVNULL)


def _oracle_rx(iq, Fs, Rs, code):
    import pirip_b200 as pb
    _, v = pb.ldpc_code(code)
    cfg = orc.demod_cfg(Fs, Rs, 2, P=8, limits=(0, Fs // 2), soft=True)
    ref = orc.fsk_demod_stream(cfg, iq)
    d = orc.Deframer(v, max_nbits=100)
    out = []
    for k in range(ref["nframes"]):
        st, pl = d.push(ref["llr"][k * 50:(k + 1) * 50])
        if st & orc.RX_BITS:
            out.append(np.packbits(pl).tobytes())
    d.close()
    return out


def test_testframe_bursts_decode_with_the_oracle_chain(tmp_path):
    _build()
    f = tmp_path / "a.u8"
    subprocess.run([os.path.join(BIN, "rpitx_fsk"), "/dev/zero", "--code", "H_256_512_4", "-r", "10000", "-s", "10000", "--testframes", "3",
                    "--bursts", "2", "--seq", "--source", "0x7", "--fs", "80000", "--ebno", "10", "-o", str(f)], check=True,
                   capture_output=True)
    iq = np.fromfile(f, np.uint8)
    got = _oracle_rx(iq, 80000, 10000, "H_256_512_4")
    assert len(got) == 6
    payload = orc.payload_data_bits(256).copy()
    for k, g in enumerate(got):
        want = payload.copy()
        want[0:8] = [0, 0, 0, 0, 0, 1, 1, 1]
        want[8:16] = [((k % 3) + 1) >> (7 - i) & 1 for i in range(8)]
        assert g == np.packbits(orc.insert_crc(want)).tobytes()


def test_packed_burst_control_input(tmp_path):
    """the byte stream tx/frame_repeater.c writes: control byte 1 / 0 / 2 + 32 packed payload bytes [R tx/frame_repeater.c:92-104]"""
    _build()
    rng = np.random.default_rng(5)
    frames = [bytes([0x02]) + rng.integers(0, 256, 31, dtype=np.uint8).tobytes() for _ in range(2)]
    stream = bytes([1]) + frames[0] + bytes([0]) + frames[1] + bytes([2]) + bytes(32)
    f = tmp_path / "b.u8"
    subprocess.run([os.path.join(BIN, "rpitx_fsk"), "-", "--code", "H_256_512_4", "-r", "10000", "-s", "10000", "--packed", "--fs", "80000",
                    "--ebno", "12", "-o", str(f)], input=stream, check=True, capture_output=True)
    got = _oracle_rx(np.fromfile(f, np.uint8), 80000, 10000, "H_256_512_4")
    assert len(got) == 2
    for g, sent in zip(got, frames):
        assert g[:30] == sent[:30]                      # the last two bytes are replaced by the CRC
        bits = np.unpackbits(np.frombuffer(g, np.uint8))
        assert np.array_equal(orc.insert_crc(bits.copy()), bits)


def test_uncoded_test_bits_through_the_oracle_demod(tmp_path):
    _build()
    bits = subprocess.run([os.path.join(BIN, "fsk_get_test_bits"), "-", "2000"], capture_output=True, check=True).stdout
    f = tmp_path / "t.u8"
    subprocess.run([os.path.join(BIN, "rpitx_fsk"), "-", "-r", "10000", "-s", "10000", "--fs", "240000", "--ebno", "12", "-o", str(f)],
                   input=bits, check=True, capture_output=True)
    ref = orc.fsk_demod_stream(orc.demod_cfg(240000, 10000, 2, P=24), np.fromfile(f, np.uint8))
    ber, n = synth.ber_vs_stream(ref["bits"])
    assert n >= 1000 and ber < 0.01
