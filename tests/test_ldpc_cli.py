"""codec2's stand-alone FEC command lines (ldpc_enc | ldpc_noise | ldpc_dec, SURVEY.md 8b / U13, BASELINE config 4) on
libpirip_b200.so.  The encoder is host code (checked here against the oracle without a GPU); the decoder runs the
sum-product kernel and is compared frame by frame with the oracle on the same LLRs."""
import math
import os
import subprocess

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "bin")
CODES = [("H_256_512_4", 512, 256), ("HRA_112_112", 224, 112)]


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s", "bin/ldpc_enc", "bin/ldpc_dec", "bin/ldpc_noise"],
                          stdout=subprocess.DEVNULL)


def _run(tool, *args, data=None):
    return subprocess.run([os.path.join(BIN, tool)] + [str(a) for a in args], input=data, capture_output=True, check=True)


@pytest.mark.parametrize("name,N,K", CODES)
def test_ldpc_enc_matches_the_oracle_encoder(name, N, K):
    import pirip_b200 as pb
    _build()
    _, v = pb.ldpc_code(name)
    # --testframes: the ofdm_rand payload, data then parity
    cw = np.frombuffer(_run("ldpc_enc", "/dev/zero", "-", "--code", name, "--testframes", 2).stdout, np.uint8)
    pay = orc.payload_data_bits(K)
    want = np.concatenate([pay, orc.ldpc_encode(v, pay)])
    assert cw.size == 2 * N and np.array_equal(cw[:N], want) and np.array_equal(cw[N:], want)
    # data from stdin, soft-decision output 1 - 2 bit
    rng = np.random.default_rng(11)
    data = rng.integers(0, 2, 3 * K, dtype=np.uint8)
    sd = np.frombuffer(_run("ldpc_enc", "-", "-", "--code", name, "--sd", data=data.tobytes()).stdout, np.float32)
    assert sd.size == 3 * N
    for f in range(3):
        d = data[f * K:(f + 1) * K]
        want = np.concatenate([d, orc.ldpc_encode(v, d)])
        assert np.array_equal(sd[f * N:(f + 1) * N], 1.0 - 2.0 * want.astype(np.float32))


def test_ldpc_noise_adds_the_requested_noise():
    _build()
    x = np.where(np.arange(20000) % 2 == 0, 1.0, -1.0).astype(np.float32)
    y = np.frombuffer(_run("ldpc_noise", "-", "-", "-3", data=x.tobytes()).stdout, np.float32)
    assert y.size == x.size
    var = float(np.var(y - x))
    assert abs(var - 0.5 * 10 ** (-0.3)) < 0.02  # No / 2 with No = -3 dB


def _sd_to_llr(sd):
    """[U mpdecode_core.c::sd_to_llr] restated with the same sequential double arithmetic"""
    n = sd.size
    s = 0.0
    for v in sd:
        s += abs(float(v))
    mean = s / n
    s = ss = 0.0
    for v in sd:
        v = float(v)
        x = v / mean - ((v > 0.0) - (v < 0.0))
        s += x
        ss += x * x
    estvar = (n * ss - s * s) / (float(n) * (n - 1))
    es = 1.0 / (2.0 * estvar + 1E-3)
    return np.array([np.float32(4.0 * es * float(v)) for v in sd], np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("name,N,K", CODES)
def test_ldpc_dec_llr_and_sd_input_against_the_oracle(name, N, K, tmp_path):
    import pirip_b200 as pb
    _build()
    l, v = pb.ldpc_code(name)
    rng = np.random.default_rng(5)
    nfr = 12
    data = rng.integers(0, 2, (nfr, K), dtype=np.uint8)
    cw = np.stack([np.concatenate([d, pb.ldpc_encode(l, d)]) for d in data]).astype(np.float64)
    ebno = np.repeat([6.0, 2.0, 1.0, 0.0], nfr // 4)  # the last frames fail now and then: iteration counts and failures must agree too
    sigma2 = 1.0 / (2 * 0.5 * 10 ** (ebno / 10))
    y = (1 - 2 * cw) + rng.standard_normal(cw.shape) * np.sqrt(sigma2)[:, None]
    llr = (2 * y / sigma2[:, None]).astype(np.float32)
    out = np.frombuffer(_run("ldpc_dec", "-", "-", "--code", name, data=llr.tobytes()).stdout, np.uint8).reshape(nfr, K)
    for f in range(nfr):
        ref, _, _ = orc.ldpc_decode(v, llr[f])
        assert np.array_equal(out[f], ref[:K])
    assert np.array_equal(out[:nfr // 4], data[:nfr // 4])  # 6 dB frames decode
    sd = y.astype(np.float32)
    out = np.frombuffer(_run("ldpc_dec", "-", "-", "--code", name, "--sd", data=sd.tobytes()).stdout, np.uint8).reshape(nfr, K)
    for f in range(nfr):
        ref, _, _ = orc.ldpc_decode(v, _sd_to_llr(sd[f]))
        assert np.array_equal(out[f], ref[:K])


@pytest.mark.gpu
def test_codec2_testframe_pipeline(tmp_path):
    """ldpc_enc --sd --testframes | ldpc_noise | ldpc_dec --sd --testframes: raw errors, no coded errors at Es/No = 1 dB"""
    _build()
    sd = _run("ldpc_enc", "/dev/zero", "-", "--sd", "--code", "H_256_512_4", "--testframes", 40).stdout
    noisy = _run("ldpc_noise", "-", "-", "-1", data=sd).stdout
    r = _run("ldpc_dec", "-", "-", "--sd", "--code", "H_256_512_4", "--testframes", data=noisy)
    err = r.stderr.decode()
    raw = [ln for ln in err.splitlines() if ln.startswith("Raw")][0].split()
    coded = [ln for ln in err.splitlines() if ln.startswith("Coded")][0].split()
    assert int(raw[2]) == 40 * 512 and int(raw[4]) > 0.03 * 40 * 512
    assert int(coded[2]) == 40 * 256 and int(coded[4]) == 0
    out = np.frombuffer(r.stdout, np.uint8).reshape(40, 256)
    assert np.array_equal(out, np.tile(orc.payload_data_bits(256), (40, 1)))
    # half-frame sync: the same stream with half a frame of junk in front still yields the frames
    junk = np.random.default_rng(2).standard_normal(256).astype(np.float32).tobytes()  # (all zeros would "decode": the zero codeword)
    r = _run("ldpc_dec", "-", "-", "--sd", "--half", "--code", "H_256_512_4", data=junk + noisy)
    out = np.frombuffer(r.stdout, np.uint8).reshape(-1, 256)
    assert out.shape[0] >= 38 and all(np.array_equal(o, orc.payload_data_bits(256)) for o in out)
