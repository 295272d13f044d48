"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module; the product package pirip_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class OrcDemodCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "Fs", "Rs", "M", "P", "Nsym", "f1_tx", "tone_spacing", "est_min", "est_max",
        "have_limits", "freq_est_type", "fmt", "soft")]


class OrcFrameStats(C.Structure):
    _fields_ = [
        ("f_est", C.c_float * 4), ("f2_est", C.c_float * 4), ("norm_rx_timing", C.c_float),
        ("ppm", C.c_float), ("SNRest", C.c_float), ("v_est", C.c_float), ("rx_sig_pow", C.c_float),
        ("rx_nse_pow", C.c_float), ("EbNodB", C.c_float), ("snr_est", C.c_float), ("foff", C.c_float),
        ("rx_timing", C.c_float), ("nin_next", C.c_int32), ("sample_offset", C.c_int32)]


STATS_DTYPE = np.dtype([
    ("f_est", np.float32, 4), ("f2_est", np.float32, 4), ("norm_rx_timing", np.float32),
    ("ppm", np.float32), ("SNRest", np.float32), ("v_est", np.float32), ("rx_sig_pow", np.float32),
    ("rx_nse_pow", np.float32), ("EbNodB", np.float32), ("snr_est", np.float32), ("foff", np.float32),
    ("rx_timing", np.float32), ("nin_next", np.int32), ("sample_offset", np.int32)])
assert STATS_DTYPE.itemsize == C.sizeof(OrcFrameStats)


class OrcLdpc(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("CodeLength", C.c_int), ("NumberParityBits", C.c_int),
                ("NumberRowsHcols", C.c_int), ("max_row_weight", C.c_int), ("max_col_weight", C.c_int),
                ("H_rows", C.POINTER(C.c_uint16)), ("H_cols", C.POINTER(C.c_uint16))]


FMT_U8, FMT_S16, FMT_F32 = 0, 1, 2
RX_TRIAL_SYNC, RX_SYNC, RX_BITS, RX_BIT_ERRORS = 1, 2, 4, 8


def build(force=False):
    """Compile the oracle with its committed Makefile (gcc -O2)."""
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if force or not os.path.exists(so) or not os.path.exists(os.path.join(ORACLE_DIR, "liboracle_libm.so")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], stdout=subprocess.DEVNULL)
    return so


_libs = {}


def lib(libm=False):
    key = "libm" if libm else "exact"
    if key in _libs:
        return _libs[key]
    build()
    L = C.CDLL(os.path.join(ORACLE_DIR, "liboracle_libm.so" if libm else "liboracle.so"))
    L.orc_fsk_demod_stream.restype = C.c_size_t
    L.orc_fsk_demod_stream.argtypes = [C.POINTER(OrcDemodCfg), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_size_t]
    L.orc_csdr_decimate_stream.restype = C.c_size_t
    L.orc_csdr_decimate_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    L.orc_firdes_filter_len.restype = C.c_int
    L.orc_firdes_filter_len.argtypes = [C.c_float]
    L.orc_firdes_lowpass_f.argtypes = [C.c_void_p, C.c_int, C.c_float]
    L.orc_fir_decimate_cc.restype = C.c_size_t
    L.orc_fir_decimate_cc.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
    L.orc_convert_u8_f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.orc_convert_f_s16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.orc_fsk_rx_filt_to_llrs.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int]
    L.orc_ldpc_encode.argtypes = [C.POINTER(OrcLdpc), C.c_void_p, C.c_void_p]
    L.orc_run_ldpc_decoder.restype = C.c_int
    L.orc_run_ldpc_decoder.argtypes = [C.POINTER(OrcLdpc), C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.orc_phi0.restype = C.c_float
    L.orc_phi0.argtypes = [C.c_float]
    L.orc_gen_crc16.restype = C.c_uint16
    L.orc_gen_crc16.argtypes = [C.c_void_p, C.c_int]
    L.orc_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.orc_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.orc_check_crc16_unpacked.restype = C.c_int
    L.orc_check_crc16_unpacked.argtypes = [C.c_void_p, C.c_int]
    L.orc_uw_bits.argtypes = [C.c_void_p]
    L.orc_generate_payload_data_bits.argtypes = [C.c_void_p, C.c_int]
    L.orc_test_bits_frame.argtypes = [C.c_void_p, C.c_int]
    L.orc_insert_crc.argtypes = [C.c_void_p, C.c_int]
    L.orc_fsk_ldpc_framer.argtypes = [C.POINTER(OrcLdpc), C.c_void_p, C.c_void_p]
    L.orc_deframer_create.restype = C.c_void_p
    L.orc_deframer_create.argtypes = [C.POINTER(OrcLdpc), C.c_int]
    L.orc_deframer_destroy.argtypes = [C.c_void_p]
    L.orc_deframer_push.restype = C.c_int
    L.orc_deframer_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_deframer_info.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_fft_alloc.restype = C.c_void_p
    L.orc_fft_alloc.argtypes = [C.c_int]
    L.orc_fft_free.argtypes = [C.c_void_p]
    L.orc_fft_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_fft_twiddles.argtypes = [C.c_void_p, C.c_void_p]
    for fn in ("orc_cosf", "orc_sinf", "orc_expf", "orc_logf", "orc_log10f"):
        getattr(L, fn).restype = C.c_float
        getattr(L, fn).argtypes = [C.c_float]
    L.orc_atan2f.restype = C.c_float
    L.orc_atan2f.argtypes = [C.c_float, C.c_float]
    L.orc_math_backend.restype = C.c_int
    _libs[key] = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def demod_cfg(Fs, Rs, M, P=8, Nsym=50, f1_tx=1200, tone_spacing=400, limits=None, mask=False, fmt=FMT_U8, soft=False):
    cfg = OrcDemodCfg(Fs=Fs, Rs=Rs, M=M, P=P, Nsym=Nsym, f1_tx=f1_tx, tone_spacing=tone_spacing,
                      est_min=limits[0] if limits else 0, est_max=limits[1] if limits else 0,
                      have_limits=1 if limits else 0, freq_est_type=1 if mask else 0, fmt=fmt, soft=1 if soft else 0)
    return cfg


def derived(cfg):
    Ts = cfg.Fs // cfg.Rs
    N = Ts * cfg.Nsym
    Nbits = cfg.Nsym * (1 if cfg.M == 2 else 2)
    return Ts, N, Nbits


def fsk_demod_stream(cfg, iq, libm=False):
    """Run the oracle's fsk_demod main loop over one capture.

    iq: uint8[2*n] / int16[2*n] / complex64[n].  Returns dict(bits, llr, rx_filt, stats, nframes)."""
    L = lib(libm)
    iq = np.ascontiguousarray(iq)
    nsamples = iq.size if cfg.fmt == FMT_F32 else iq.size // 2
    Ts, N, Nbits = derived(cfg)
    max_frames = nsamples // (N - Ts // 4) + 1
    bits = np.zeros(max_frames * Nbits, np.uint8)
    llr = np.zeros(max_frames * Nbits, np.float32) if cfg.soft else None
    rx_filt = np.zeros(max_frames * cfg.M * cfg.Nsym, np.float32)
    stats = np.zeros(max_frames, STATS_DTYPE)
    nf = L.orc_fsk_demod_stream(C.byref(cfg), _ptr(iq), nsamples, _ptr(bits), _ptr(llr), _ptr(rx_filt), _ptr(stats),
                                max_frames)
    return dict(nframes=nf, bits=bits[:nf * Nbits], llr=None if llr is None else llr[:nf * Nbits],
                rx_filt=rx_filt[:nf * cfg.M * cfg.Nsym], stats=stats[:nf])


class OrcComp(C.Structure):
    _fields_ = [("real", C.c_float), ("imag", C.c_float)]


class OrcFsk(C.Structure):
    """struct orc_fsk of oracle/oracle.h ([U fsk.h::struct FSK])."""
    _fields_ = [("Ndft", C.c_int), ("Fs", C.c_int), ("N", C.c_int), ("Rs", C.c_int), ("Ts", C.c_int), ("Nmem", C.c_int),
                ("P", C.c_int), ("Nsym", C.c_int), ("Nbits", C.c_int), ("f1_tx", C.c_int), ("tone_spacing", C.c_int),
                ("mode", C.c_int), ("tc", C.c_float), ("est_min", C.c_int), ("est_max", C.c_int), ("est_space", C.c_int),
                ("hann_table", C.c_void_p), ("Sf", C.POINTER(C.c_float)), ("phi_c", OrcComp * 4), ("f_dc", C.c_void_p),
                ("fft_cfg", C.c_void_p), ("norm_rx_timing", C.c_float), ("tx_phase_c", OrcComp), ("EbNodB", C.c_float),
                ("f_est", C.c_float * 4), ("f2_est", C.c_float * 4), ("freq_est_type", C.c_int), ("ppm", C.c_float),
                ("SNRest", C.c_float), ("v_est", C.c_float), ("rx_sig_pow", C.c_float), ("rx_nse_pow", C.c_float),
                ("nin", C.c_int), ("burst_mode", C.c_int), ("lock_nin", C.c_int), ("stats_snr_est", C.c_float),
                ("stats_foff", C.c_float), ("stats_rx_timing", C.c_float), ("stats_clock_offset", C.c_float),
                ("normalise_eye", C.c_int)]


class FskStream:
    """The oracle's per-stream demodulator object, call for call like codec2's struct FSK API (the mirror of
    pirip_b200.FskStream): create_hbr / nin / demod / demod_sd / clear_estimators / enable_burst_mode."""

    def __init__(self, Fs, Rs, M, P=8, Nsym=50, f1_tx=1200, tone_spacing=400, limits=None, mask=False):
        L = lib()
        L.orc_fsk_create_hbr.restype = C.POINTER(OrcFsk)
        L.orc_fsk_create_hbr.argtypes = [C.c_int] * 7
        L.orc_fsk_destroy.argtypes = [C.POINTER(OrcFsk)]
        L.orc_fsk_nin.restype = C.c_uint32
        L.orc_fsk_nin.argtypes = [C.POINTER(OrcFsk)]
        L.orc_fsk_demod.argtypes = [C.POINTER(OrcFsk), C.c_void_p, C.c_void_p]
        L.orc_fsk_demod_sd.argtypes = [C.POINTER(OrcFsk), C.c_void_p, C.c_void_p]
        L.orc_fsk_set_freq_est_limits.argtypes = [C.POINTER(OrcFsk), C.c_int, C.c_int]
        L.orc_fsk_set_freq_est_alg.argtypes = [C.POINTER(OrcFsk), C.c_int]
        L.orc_fsk_clear_estimators.argtypes = [C.POINTER(OrcFsk)]
        L.orc_fsk_enable_burst_mode.argtypes = [C.POINTER(OrcFsk)]
        self.L = L
        self.fsk = L.orc_fsk_create_hbr(Fs, Rs, M, P, Nsym, f1_tx, tone_spacing)
        if limits:
            L.orc_fsk_set_freq_est_limits(self.fsk, limits[0], limits[1])
        if mask:
            L.orc_fsk_set_freq_est_alg(self.fsk, 1)

    @property
    def s(self):
        return self.fsk.contents

    def nin(self):
        return int(self.L.orc_fsk_nin(self.fsk))

    def demod(self, samples):
        samples = np.ascontiguousarray(samples, np.complex64)
        assert samples.size == self.nin()
        bits = np.zeros(self.s.Nbits, np.uint8)
        self.L.orc_fsk_demod(self.fsk, _ptr(bits), _ptr(samples))
        return bits

    def demod_sd(self, samples):
        samples = np.ascontiguousarray(samples, np.complex64)
        filt = np.zeros(self.s.mode * self.s.Nsym, np.float32)
        self.L.orc_fsk_demod_sd(self.fsk, _ptr(filt), _ptr(samples))
        return filt

    def clear_estimators(self):
        self.L.orc_fsk_clear_estimators(self.fsk)

    def enable_burst_mode(self):
        self.L.orc_fsk_enable_burst_mode(self.fsk)

    def close(self):
        if self.fsk:
            self.L.orc_fsk_destroy(self.fsk)
            self.fsk = None


def csdr_decimate_stream(iq_u8, decimation, transition_bw=0.05, libm=False):
    L = lib(libm)
    iq_u8 = np.ascontiguousarray(iq_u8, np.uint8)
    n = iq_u8.size // 2
    out_f = np.zeros(n // decimation + 1, np.complex64)
    out_s = np.zeros(2 * (n // decimation + 1), np.int16)
    nout = L.orc_csdr_decimate_stream(_ptr(iq_u8), n, decimation, transition_bw, _ptr(out_f), _ptr(out_s))
    return out_f[:nout], out_s[:2 * nout]


def firdes_lowpass(decimation, transition_bw=0.05):
    L = lib()
    n = L.orc_firdes_filter_len(transition_bw)
    taps = np.zeros(n, np.float32)
    L.orc_firdes_lowpass_f(_ptr(taps), n, np.float32(0.5 / np.float32(decimation)))
    return taps


def make_ldpc(code, max_iter=15):
    """code: dict with CodeLength, NumberParityBits, NumberRowsHcols, max_row_weight, max_col_weight,
    H_rows (uint16), H_cols (uint16).  Returns (OrcLdpc, keepalive)."""
    hr = np.ascontiguousarray(code["H_rows"], np.uint16)
    hc = np.ascontiguousarray(code["H_cols"], np.uint16)
    l = OrcLdpc(max_iter=max_iter, CodeLength=code["CodeLength"], NumberParityBits=code["NumberParityBits"],
                NumberRowsHcols=code["NumberRowsHcols"], max_row_weight=code["max_row_weight"],
                max_col_weight=code["max_col_weight"],
                H_rows=hr.ctypes.data_as(C.POINTER(C.c_uint16)), H_cols=hc.ctypes.data_as(C.POINTER(C.c_uint16)))
    return l, (hr, hc)


def ldpc_encode(code, ibits):
    L = lib()
    l, keep = make_ldpc(code)
    ibits = np.ascontiguousarray(ibits, np.uint8)
    p = np.zeros(code["NumberParityBits"], np.uint8)
    L.orc_ldpc_encode(C.byref(l), _ptr(ibits), _ptr(p))
    return p


def ldpc_decode(code, llr, max_iter=15, libm=False):
    L = lib(libm)
    l, keep = make_ldpc(code, max_iter)
    llr = np.ascontiguousarray(llr, np.float32)
    out = np.zeros(code["CodeLength"], np.uint8)
    pcc = C.c_int(0)
    it = L.orc_run_ldpc_decoder(C.byref(l), _ptr(out), _ptr(llr), C.byref(pcc))
    return out, it, pcc.value


def rx_filt_to_llrs(rx_filt, v_est, SNRest, M, nsyms, libm=False):
    L = lib(libm)
    rx_filt = np.ascontiguousarray(rx_filt, np.float32)
    llr = np.zeros(nsyms * (1 if M == 2 else 2), np.float32)
    L.orc_fsk_rx_filt_to_llrs(_ptr(llr), _ptr(rx_filt), np.float32(v_est), np.float32(SNRest), M, nsyms)
    return llr


def crc16(data):
    b = np.frombuffer(bytes(data), np.uint8)
    return lib().orc_gen_crc16(_ptr(b), b.size)


def uw_bits():
    u = np.zeros(32, np.uint8)
    lib().orc_uw_bits(_ptr(u))
    return u


def payload_data_bits(n):
    b = np.zeros(n, np.uint8)
    lib().orc_generate_payload_data_bits(_ptr(b), n)
    return b


def test_bits_frame(n=100):
    b = np.zeros(n, np.uint8)
    lib().orc_test_bits_frame(_ptr(b), n)
    return b


test_bits_frame.__test__ = False


def insert_crc(bits):
    b = np.ascontiguousarray(bits, np.uint8).copy()
    lib().orc_insert_crc(_ptr(b), b.size)
    return b


def fsk_ldpc_frame(code, payload):
    L = lib()
    l, keep = make_ldpc(code)
    payload = np.ascontiguousarray(payload, np.uint8)
    frame = np.zeros(32 + code["CodeLength"], np.uint8)
    L.orc_fsk_ldpc_framer(C.byref(l), _ptr(frame), _ptr(payload))
    return frame


class Deframer:
    def __init__(self, code, max_nbits=200, max_iter=15, libm=False):
        self.L = lib(libm)
        self.l, self.keep = make_ldpc(code, max_iter)
        self.K = code["CodeLength"] - code["NumberParityBits"]
        self.h = self.L.orc_deframer_create(C.byref(self.l), max_nbits)

    def push(self, llr):
        llr = np.ascontiguousarray(llr, np.float32)
        out = np.zeros(self.K, np.uint8)
        st = self.L.orc_deframer_push(self.h, _ptr(llr), llr.size, _ptr(out))
        return st, out

    def info(self):
        """state, uw_loc, uw_err, iter, pcc, nbits queued after the last push (the `rtl_fsk -v` log fields)"""
        a = np.zeros(6, np.int32)
        self.L.orc_deframer_info(self.h, _ptr(a))
        return a

    def close(self):
        if self.h:
            self.L.orc_deframer_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def fft(x, libm=False):
    L = lib(libm)
    x = np.ascontiguousarray(x, np.complex64)
    cfg = L.orc_fft_alloc(x.size)
    out = np.zeros_like(x)
    L.orc_fft_run(cfg, _ptr(x), _ptr(out))
    L.orc_fft_free(cfg)
    return out
