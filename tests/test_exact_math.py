"""exact_math.h: the bit-reproducible transcendentals must be within 1 ulp of glibc, and the
oracle built with them must agree with the oracle built with glibc's own functions."""
import ctypes as C

import numpy as np

import orc
import synth


def ulp_diff(a, b):
    ai = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def _call(fn, xs):
    return np.array([fn(C.c_float(float(x))) for x in xs], np.float32)


def test_backends():
    assert orc.lib().orc_math_backend() == 0
    assert orc.lib(libm=True).orc_math_backend() == 1


def test_sincos_within_one_ulp_of_libm():
    g = synth.rng(1)
    xs = np.concatenate([g.uniform(-np.pi, np.pi, 4000), g.uniform(-50, 50, 1000), [0.0, 1e-8, -1e-8]]).astype(np.float32)
    E, R = orc.lib(), orc.lib(libm=True)
    for name in ("orc_cosf", "orc_sinf"):
        d = ulp_diff(_call(getattr(E, name), xs), _call(getattr(R, name), xs))
        assert d.max() <= 1, (name, d.max())
        assert (d == 0).mean() > 0.97  # glibc is faithfully, not correctly, rounded


def test_atan2_within_one_ulp_of_libm():
    g = synth.rng(2)
    y = g.standard_normal(5000).astype(np.float32)
    x = g.standard_normal(5000).astype(np.float32)
    E, R = orc.lib(), orc.lib(libm=True)
    a = np.array([E.orc_atan2f(C.c_float(float(p)), C.c_float(float(q))) for p, q in zip(y, x)], np.float32)
    b = np.array([R.orc_atan2f(C.c_float(float(p)), C.c_float(float(q))) for p, q in zip(y, x)], np.float32)
    assert ulp_diff(a, b).max() <= 1
    # quadrant / axis cases
    for p, q, want in [(0, 1, 0.0), (1, 0, np.pi / 2), (0, -1, np.pi), (-1, 0, -np.pi / 2), (0, 0, 0.0)]:
        assert abs(E.orc_atan2f(C.c_float(p), C.c_float(q)) - want) < 1e-6


def test_exp_log_within_one_ulp_of_libm():
    g = synth.rng(3)
    E, R = orc.lib(), orc.lib(libm=True)
    xs = g.uniform(-20, 20, 4000).astype(np.float32)
    assert ulp_diff(_call(E.orc_expf, xs), _call(R.orc_expf, xs)).max() <= 1
    ps = np.exp(g.uniform(-30, 30, 4000)).astype(np.float32)
    assert ulp_diff(_call(E.orc_logf, ps), _call(R.orc_logf, ps)).max() <= 1
    assert ulp_diff(_call(E.orc_log10f, ps), _call(R.orc_log10f, ps)).max() <= 1


def test_oracle_with_libm_agrees_with_exact_math_oracle():
    """The substitution of exact_math for libm is itself pinned: same hard bits, same nin
    sequence, floats within a few ulp-scale relative error, on noisy captures."""
    Fs, Rs = 240000, 10000
    for M, P, soft in ((2, 24, False), (4, 8, True)):
        cfg = orc.demod_cfg(Fs, Rs, M, P=P, limits=(500, 45000), soft=soft)
        for c, eb in enumerate([None, 9.0, 4.0]):
            u8, _ = synth.capture_u8(40 + c, 6000 * (2 if M == 4 else 1), M, Fs, Rs, 10000, 10000, ebno_db=eb)
            a = orc.fsk_demod_stream(cfg, u8)
            b = orc.fsk_demod_stream(cfg, u8, libm=True)
            assert a["nframes"] == b["nframes"]
            assert np.array_equal(a["bits"], b["bits"])
            assert np.array_equal(a["stats"]["nin_next"], b["stats"]["nin_next"])
            assert np.array_equal(a["stats"]["f_est"], b["stats"]["f_est"])
            # a 1-ulp difference in the oscillator step e^{j dphi} is integrated over the whole capture
            assert np.allclose(a["stats"]["norm_rx_timing"], b["stats"]["norm_rx_timing"], atol=1e-4)
            assert np.allclose(a["rx_filt"], b["rx_filt"], rtol=1e-3, atol=1e-4)
            if soft:
                assert np.allclose(a["llr"], b["llr"], rtol=5e-3, atol=5e-2)
