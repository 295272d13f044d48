"""ctypes loader for libpirip_b200.so and Python mirrors of the C-ABI structs."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FMT_U8, FMT_S16, FMT_F32 = 0, 1, 2
MODE_EXACT, MODE_FAST = 0, 1


class PiripError(RuntimeError):
    pass


def lib_path():
    # PIRIP_B200_LIB: another build of the same library (A/B timing of two source versions, tools/ab_k2.py)
    return os.environ.get("PIRIP_B200_LIB") or os.path.join(_HERE, "libpirip_b200.so")


_lib = None


class FskCfg(C.Structure):
    """struct pirip_fsk_cfg (include/pirip_fsk.h)"""
    _fields_ = [(n, C.c_int) for n in (
        "Fs", "Rs", "M", "P", "Nsym", "f1_tx", "tone_spacing", "est_min", "est_max", "have_limits",
        "freq_est_type", "fmt", "soft")]


FRAME_STATS_DTYPE = np.dtype([
    ("f_est", np.float32, 4), ("f2_est", np.float32, 4), ("norm_rx_timing", np.float32),
    ("ppm", np.float32), ("SNRest", np.float32), ("v_est", np.float32), ("rx_sig_pow", np.float32),
    ("rx_nse_pow", np.float32), ("EbNodB", np.float32), ("snr_est", np.float32), ("foff", np.float32),
    ("rx_timing", np.float32), ("nin_next", np.int32), ("sample_offset", np.int32)])


def lib():
    """Load the shared library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise PiripError(
            f"{path} not found: build it with `make -C pirip_b200/csrc` (or __graft_entry__.build()); "
            "there is no CPU fallback")
    L = C.CDLL(path)
    L.pirip_fsk_plan_create.restype = C.c_void_p
    L.pirip_fsk_plan_create.argtypes = [C.POINTER(FskCfg), C.c_int]
    L.pirip_fsk_plan_destroy.argtypes = [C.c_void_p]
    L.pirip_fsk_plan_nbits.restype = C.c_int
    L.pirip_fsk_plan_nbits.argtypes = [C.c_void_p]
    L.pirip_fsk_plan_nfilt.restype = C.c_int
    L.pirip_fsk_plan_nfilt.argtypes = [C.c_void_p]
    L.pirip_fsk_plan_max_frames.restype = C.c_size_t
    L.pirip_fsk_plan_max_frames.argtypes = [C.c_void_p, C.c_size_t]
    L.pirip_fsk_demod_batch_dev.restype = C.c_int
    L.pirip_fsk_demod_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.pirip_fsk_demod_batch.restype = C.c_int
    L.pirip_fsk_demod_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.pirip_fsk_plan_set_mode.restype = C.c_int
    L.pirip_fsk_plan_set_mode.argtypes = [C.c_void_p, C.c_int]
    L.pirip_fsk_plan_mode.restype = C.c_int
    L.pirip_fsk_plan_mode.argtypes = [C.c_void_p]
    L.pirip_fsk_plan_flagged.restype = C.c_long
    L.pirip_fsk_plan_flagged.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.pirip_fsk_plan_replays.restype = C.c_int
    L.pirip_fsk_plan_replays.argtypes = [C.c_void_p, C.c_void_p]
    L.pirip_fsk_plan_kernel.restype = C.c_char_p
    L.pirip_fsk_plan_kernel.argtypes = [C.c_void_p]
    L.pirip_pack_bits_dev.restype = C.c_int
    L.pirip_pack_bits_dev.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.pirip_kernel_launches.restype = C.c_uint64
    L.pirip_last_error.restype = C.c_char_p
    L.pirip_version.restype = C.c_char_p
    _lib = L
    return L


def version():
    return lib().pirip_version().decode()


def kernel_launches():
    return int(lib().pirip_kernel_launches())


def pack_bits_dev(d_bits, nbits, d_packed, stream=None):
    """Device pointers in, MSB-first packing (freedv_pack order); asynchronous on `stream`."""
    rc = lib().pirip_pack_bits_dev(d_bits, nbits, d_packed, stream)
    if rc != 0:
        raise _err("pirip_pack_bits_dev", rc)


def _err(what, rc=None):
    msg = lib().pirip_last_error().decode()
    return PiripError(f"{what} failed" + (f" (rc={rc})" if rc is not None else "") + (f": {msg}" if msg else ""))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class FskPlan:
    """Python mirror of the batched demodulator interface.

    Argument meaning follows the fsk_demod CLI ([R README.md:105,109]):
    `fsk_demod [-c|-d] [-p P] [--fsk_lower Hz] [--fsk_upper Hz] [--mask spacing] [-s] M Fs Rs in out`.
    """

    def __init__(self, Fs, Rs, M, P=8, Nsym=50, f1_tx=1200, tone_spacing=400, limits=None, mask=False,
                 fmt=FMT_U8, soft=False, device=0):
        self.cfg = FskCfg(Fs=Fs, Rs=Rs, M=M, P=P, Nsym=Nsym, f1_tx=f1_tx, tone_spacing=tone_spacing,
                          est_min=limits[0] if limits else 0, est_max=limits[1] if limits else 0,
                          have_limits=1 if limits else 0, freq_est_type=1 if mask else 0, fmt=fmt,
                          soft=1 if soft else 0)
        self._h = lib().pirip_fsk_plan_create(C.byref(self.cfg), device)
        if not self._h:
            raise _err("pirip_fsk_plan_create")
        self.device = device
        self.nbits = lib().pirip_fsk_plan_nbits(self._h)
        self.nfilt = lib().pirip_fsk_plan_nfilt(self._h)
        self.bytes_per_sample = {FMT_U8: 2, FMT_S16: 4, FMT_F32: 8}[fmt]

    def max_frames(self, samples_per_capture):
        return int(lib().pirip_fsk_plan_max_frames(self._h, samples_per_capture))

    # ---- numerics mode (include/pirip_fsk.h: PIRIP_FSK_MODE_*) ------------------------------
    def set_mode(self, mode):
        """'fast' (tolerance kernel + exact re-run of flagged captures) or 'exact'."""
        rc = lib().pirip_fsk_plan_set_mode(self._h, {"exact": MODE_EXACT, "fast": MODE_FAST}[mode])
        if rc != 0:
            raise _err("pirip_fsk_plan_set_mode", rc)

    @property
    def mode(self):
        return "fast" if lib().pirip_fsk_plan_mode(self._h) == 1 else "exact"

    def kernel(self):
        return lib().pirip_fsk_plan_kernel(self._h).decode()

    def replays(self):
        """(replayed, failed) exact in-kernel frame replays since the plan was created."""
        out = np.zeros(2, np.int32)
        rc = lib().pirip_fsk_plan_replays(self._h, _ptr(out))
        if rc != 0:
            raise _err("pirip_fsk_plan_replays", rc)
        return int(out[0]), int(out[1])

    def ppm_near_ties(self):
        """frames whose ppm update decision lay inside the estimates' error of its threshold (taken from the fast values)"""
        L = lib()
        L.pirip_fsk_plan_ppm_near_ties.restype = C.c_long
        L.pirip_fsk_plan_ppm_near_ties.argtypes = [C.c_void_p]
        return int(L.pirip_fsk_plan_ppm_near_ties(self._h))

    def flagged(self, n=0):
        """(count, flags[n]) of the last batch call: captures the guards sent to the exact kernel."""
        flags = np.zeros(n, np.int32) if n else None
        cnt = lib().pirip_fsk_plan_flagged(self._h, _ptr(flags), n)
        if cnt < 0:
            raise _err("pirip_fsk_plan_flagged", cnt)
        return int(cnt), flags

    def close(self):
        if getattr(self, "_h", None):
            lib().pirip_fsk_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host buffers in, host buffers out (copies inside) -------------------------------
    def demod_batch(self, iq, want_llr=False, want_rx_filt=False, want_stats=False):
        """iq: array [n_captures, ...] of one capture per row (uint8 2*n, int16 2*n or complex64 n).
        Returns dict(bits[n_captures, max_frames*nbits], nframes[n_captures], llr, rx_filt, stats)."""
        iq = np.ascontiguousarray(iq)
        assert iq.ndim == 2
        ncap = iq.shape[0]
        spc = iq.shape[1] * iq.dtype.itemsize // self.bytes_per_sample
        mf = self.max_frames(spc)
        bits = np.zeros((ncap, mf * self.nbits), np.uint8)
        nframes = np.zeros(ncap, np.int32)
        llr = np.zeros((ncap, mf * self.nbits), np.float32) if want_llr else None
        filt = np.zeros((ncap, mf * self.nfilt), np.float32) if want_rx_filt else None
        stats = np.zeros((ncap, mf), FRAME_STATS_DTYPE) if want_stats else None
        rc = lib().pirip_fsk_demod_batch(self._h, _ptr(iq), ncap, spc, _ptr(bits), _ptr(llr), _ptr(filt),
                                         _ptr(stats), _ptr(nframes), mf)
        if rc != 0:
            raise _err("pirip_fsk_demod_batch", rc)
        return dict(bits=bits, nframes=nframes, llr=llr, rx_filt=filt, stats=stats, max_frames=mf)

    # ---- device-resident (raw device pointers, e.g. torch tensor .data_ptr()) ----------------
    def demod_batch_dev(self, d_iq, n_captures, samples_per_capture, stride_bytes, d_bits, d_nframes, max_frames,
                        d_llr=None, d_rx_filt=None, d_stats=None, stream=None):
        rc = lib().pirip_fsk_demod_batch_dev(self._h, d_iq, n_captures, samples_per_capture, stride_bytes, d_bits,
                                             d_llr, d_rx_filt, d_stats, d_nframes, max_frames, stream)
        if rc != 0:
            raise _err("pirip_fsk_demod_batch_dev", rc)


def demod_batch_multi(plans, iq, want_llr=False):
    """pirip_fsk_demod_batch_multi: one host process, one plan per GPU; the captures (rows of iq) are cut into contiguous blocks,
    every device takes one, the bit streams come back in capture order."""
    iq = np.ascontiguousarray(iq)
    ncap = iq.shape[0]
    p0 = plans[0]
    spc = iq.shape[1] * iq.dtype.itemsize // p0.bytes_per_sample
    mf = p0.max_frames(spc)
    bits = np.zeros((ncap, mf * p0.nbits), np.uint8)
    nframes = np.zeros(ncap, np.int32)
    llr = np.zeros((ncap, mf * p0.nbits), np.float32) if want_llr else None
    L = lib()
    L.pirip_fsk_demod_batch_multi.restype = C.c_int
    L.pirip_fsk_demod_batch_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_size_t]
    arr = (C.c_void_p * len(plans))(*[p._h for p in plans])
    rc = L.pirip_fsk_demod_batch_multi(arr, len(plans), _ptr(iq), ncap, spc, _ptr(bits), _ptr(llr), _ptr(nframes), mf)
    if rc != 0:
        raise _err("pirip_fsk_demod_batch_multi", rc)
    return dict(bits=bits, nframes=nframes, llr=llr, max_frames=mf)


# ---- codec2 `struct FSK` API (include/pirip_fsk.h layer 1) -------------------------------------

class COMP(C.Structure):
    _fields_ = [("real", C.c_float), ("imag", C.c_float)]


class FSK(C.Structure):
    """struct FSK as declared in include/pirip_fsk.h (codec2 fsk.h field names)."""
    _fields_ = [
        ("Ndft", C.c_int), ("Fs", C.c_int), ("N", C.c_int), ("Rs", C.c_int), ("Ts", C.c_int), ("Nmem", C.c_int),
        ("P", C.c_int), ("Nsym", C.c_int), ("Nbits", C.c_int), ("f1_tx", C.c_int), ("tone_spacing", C.c_int),
        ("mode", C.c_int), ("tc", C.c_float), ("est_min", C.c_int), ("est_max", C.c_int), ("est_space", C.c_int),
        ("hann_table", C.POINTER(C.c_float)), ("Sf", C.POINTER(C.c_float)), ("phi_c", COMP * 4),
        ("f_dc", C.c_void_p), ("fft_cfg", C.c_void_p), ("norm_rx_timing", C.c_float), ("tx_phase_c", COMP),
        ("EbNodB", C.c_float), ("f_est", C.c_float * 4), ("f2_est", C.c_float * 4), ("freq_est_type", C.c_int),
        ("ppm", C.c_float), ("SNRest", C.c_float), ("v_est", C.c_float), ("rx_sig_pow", C.c_float),
        ("rx_nse_pow", C.c_float), ("nin", C.c_int), ("burst_mode", C.c_int), ("lock_nin", C.c_int),
        ("stats", C.c_void_p), ("normalise_eye", C.c_int), ("dev", C.c_void_p)]


class FskStream:
    """Per-stream demodulator through codec2's own entry points: fsk_create_hbr / fsk_nin /
    fsk_demod / fsk_demod_sd / fsk_destroy, exactly as fsk_demod.c and rtl_fsk call them."""

    def __init__(self, Fs, Rs, M, P=8, Nsym=50, f1_tx=1200, tone_spacing=400, limits=None, mask=False):
        L = lib()
        L.fsk_create_hbr.restype = C.POINTER(FSK)
        L.fsk_create_hbr.argtypes = [C.c_int] * 7
        L.fsk_destroy.argtypes = [C.POINTER(FSK)]
        L.fsk_nin.restype = C.c_uint32
        L.fsk_nin.argtypes = [C.POINTER(FSK)]
        L.fsk_demod.argtypes = [C.POINTER(FSK), C.c_void_p, C.c_void_p]
        L.fsk_demod_sd.argtypes = [C.POINTER(FSK), C.c_void_p, C.c_void_p]
        L.fsk_set_freq_est_limits.argtypes = [C.POINTER(FSK), C.c_int, C.c_int]
        L.fsk_set_freq_est_alg.argtypes = [C.POINTER(FSK), C.c_int]
        self.L = L
        self.fsk = L.fsk_create_hbr(Fs, Rs, M, P, Nsym, f1_tx, tone_spacing)
        if limits:
            L.fsk_set_freq_est_limits(self.fsk, limits[0], limits[1])
        if mask:
            L.fsk_set_freq_est_alg(self.fsk, 1)

    @property
    def s(self):
        return self.fsk.contents

    def nin(self):
        return int(self.L.fsk_nin(self.fsk))

    def demod(self, samples):
        """samples: complex64[nin] -> uint8[Nbits]"""
        samples = np.ascontiguousarray(samples, np.complex64)
        assert samples.size == self.nin()
        bits = np.zeros(self.s.Nbits, np.uint8)
        self.L.fsk_demod(self.fsk, _ptr(bits), _ptr(samples))
        return bits

    def demod_sd(self, samples):
        samples = np.ascontiguousarray(samples, np.complex64)
        assert samples.size == self.nin()
        filt = np.zeros(self.s.mode * self.s.Nsym, np.float32)
        self.L.fsk_demod_sd(self.fsk, _ptr(filt), _ptr(samples))
        return filt

    def demod_frames(self, samples, max_frames):
        """pirip_fsk_demod_frames: every whole frame of `samples` (at most max_frames) in one launch.
        Returns (bits [nf, Nbits], rx_filt [nf, mode * Nsym], stats [nf], consumed)."""
        samples = np.ascontiguousarray(samples, np.complex64)
        L = self.L
        L.pirip_fsk_demod_frames.restype = C.c_int
        L.pirip_fsk_demod_frames.argtypes = [C.POINTER(FSK), C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_size_t)]
        bits = np.zeros((max_frames, self.s.Nbits), np.uint8)
        filt = np.zeros((max_frames, self.s.mode * self.s.Nsym), np.float32)
        stats = np.zeros(max_frames, FRAME_STATS_DTYPE)
        used = C.c_size_t(0)
        nf = L.pirip_fsk_demod_frames(self.fsk, _ptr(samples), samples.size, max_frames, _ptr(bits), _ptr(filt), _ptr(stats), C.byref(used))
        return bits[:nf], filt[:nf], stats[:nf], int(used.value)

    def clear_estimators(self):
        self.L.fsk_clear_estimators.argtypes = [C.POINTER(FSK)]
        self.L.fsk_clear_estimators(self.fsk)

    def enable_burst_mode(self):
        self.L.fsk_enable_burst_mode.argtypes = [C.POINTER(FSK)]
        self.L.fsk_enable_burst_mode(self.fsk)

    def mod_ext_vco(self, bits):
        """fsk_mod_ext_vco: per-sample tone frequency for an external VCO."""
        bits = np.ascontiguousarray(bits, np.uint8)
        self.L.fsk_mod_ext_vco.argtypes = [C.POINTER(FSK), C.c_void_p, C.c_void_p, C.c_int]
        out = np.zeros(bits.size // (self.s.mode >> 1) * self.s.Ts, np.float32)
        self.L.fsk_mod_ext_vco(self.fsk, _ptr(out), _ptr(bits), bits.size)
        return out

    def close(self):
        if self.fsk:
            self.L.fsk_destroy(self.fsk)
            self.fsk = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- csdr stage (include/pirip_csdr.h) ------------------------------------------------------------

class DecimPlan:
    """`csdr convert_u8_f | csdr fir_decimate_cc D [transition_bw] | csdr convert_f_s16` as one
    fused GPU kernel over a batch of u8 IQ streams ([R README.md:109])."""

    def __init__(self, decimation, transition_bw=0.05, device=0):
        L = lib()
        L.firdes_filter_len.restype = C.c_int
        L.firdes_filter_len.argtypes = [C.c_float]
        L.firdes_lowpass_f.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
        L.pirip_decim_plan_create.restype = C.c_void_p
        L.pirip_decim_plan_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.pirip_decim_plan_destroy.argtypes = [C.c_void_p]
        L.pirip_decim_out_count.restype = C.c_size_t
        L.pirip_decim_out_count.argtypes = [C.c_void_p, C.c_size_t]
        L.pirip_decim_u8_batch.restype = C.c_int
        L.pirip_decim_u8_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        L.pirip_decim_u8_batch_dev.restype = C.c_int
        L.pirip_decim_u8_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                               C.c_void_p, C.c_size_t, C.c_void_p]
        self.decimation = decimation
        n = L.firdes_filter_len(transition_bw)
        self.taps = np.zeros(n, np.float32)
        L.firdes_lowpass_f(_ptr(self.taps), n, np.float32(0.5 / np.float32(decimation)), 0)
        self._h = L.pirip_decim_plan_create(decimation, _ptr(self.taps), n, device)
        if not self._h:
            raise _err("pirip_decim_plan_create")

    def out_count(self, n):
        return int(lib().pirip_decim_out_count(self._h, n))

    def run(self, iq_u8, want_f32=True, want_s16=True):
        iq_u8 = np.ascontiguousarray(iq_u8, np.uint8)
        assert iq_u8.ndim == 2
        ns, n = iq_u8.shape[0], iq_u8.shape[1] // 2
        nout = self.out_count(n)
        f = np.zeros((ns, nout), np.complex64) if want_f32 else None
        s = np.zeros((ns, 2 * nout), np.int16) if want_s16 else None
        rc = lib().pirip_decim_u8_batch(self._h, _ptr(iq_u8), ns, n, _ptr(f), _ptr(s))
        if rc != 0:
            raise _err("pirip_decim_u8_batch", rc)
        return f, s

    def run_dev(self, d_in, n_streams, n, in_stride_bytes, d_out_f32, d_out_s16, out_stride, stream=None):
        rc = lib().pirip_decim_u8_batch_dev(self._h, d_in, n_streams, n, in_stride_bytes, d_out_f32, d_out_s16, out_stride,
                                            stream)
        if rc != 0:
            raise _err("pirip_decim_u8_batch_dev", rc)

    def close(self):
        if getattr(self, "_h", None):
            lib().pirip_decim_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- FEC / framing stage (include/pirip_ldpc.h) ---------------------------------------------------

class LDPC(C.Structure):
    """struct LDPC (codec2 mpdecode_core.h layout)"""
    _fields_ = [("name", C.c_char * 32), ("max_iter", C.c_int), ("dec_type", C.c_int), ("q_scale_factor", C.c_float),
                ("r_scale_factor", C.c_float), ("CodeLength", C.c_int), ("NumberParityBits", C.c_int),
                ("NumberRowsHcols", C.c_int), ("max_row_weight", C.c_int), ("max_col_weight", C.c_int),
                ("H_rows", C.POINTER(C.c_uint16)), ("H_cols", C.POINTER(C.c_uint16)),
                ("ldpc_data_bits_per_frame", C.c_int), ("ldpc_coded_bits_per_frame", C.c_int),
                ("protection_mode", C.c_int), ("data_bits_per_frame", C.c_int), ("coded_bits_per_frame", C.c_int)]


class FreedvAdvanced(C.Structure):
    _fields_ = [("interleave_frames", C.c_int), ("M", C.c_int), ("Rs", C.c_int), ("Fs", C.c_int),
                ("first_tone", C.c_int), ("tone_spacing", C.c_int), ("codename", C.c_char_p)]


FREEDV_MODE_FSK_LDPC = 9
RX_TRIAL_SYNC, RX_SYNC, RX_BITS, RX_BIT_ERRORS = 1, 2, 4, 8


def _ldpc_protos():
    L = lib()
    if getattr(L, "_ldpc_ready", False):
        return L
    L.ldpc_codes_find.restype = C.c_int
    L.ldpc_codes_find.argtypes = [C.c_char_p]
    L.ldpc_codes_setup.argtypes = [C.POINTER(LDPC), C.c_char_p]
    L.ldpc_codes_num.restype = C.c_int
    L.ldpc_encode_frame.argtypes = [C.POINTER(LDPC), C.c_void_p, C.c_void_p]
    L.run_ldpc_decoder.restype = C.c_int
    L.run_ldpc_decoder.argtypes = [C.POINTER(LDPC), C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.fsk_rx_filt_to_llrs.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int]
    L.pirip_ldpc_plan_create.restype = C.c_void_p
    L.pirip_ldpc_plan_create.argtypes = [C.POINTER(LDPC), C.c_int]
    L.pirip_ldpc_plan_destroy.argtypes = [C.c_void_p]
    L.pirip_ldpc_decode_batch.restype = C.c_int
    L.pirip_ldpc_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pirip_ldpc_decode_batch_dev.restype = C.c_int
    L.pirip_ldpc_decode_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]
    L.freedv_open_advanced.restype = C.c_void_p
    L.freedv_open_advanced.argtypes = [C.c_int, C.POINTER(FreedvAdvanced)]
    L.freedv_close.argtypes = [C.c_void_p]
    L.freedv_get_bits_per_modem_frame.restype = C.c_int
    L.freedv_get_bits_per_modem_frame.argtypes = [C.c_void_p]
    L.freedv_tx_fsk_ldpc_bits_per_frame.restype = C.c_int
    L.freedv_tx_fsk_ldpc_bits_per_frame.argtypes = [C.c_void_p]
    L.freedv_tx_fsk_ldpc_framer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.freedv_gen_crc16.restype = C.c_ushort
    L.freedv_gen_crc16.argtypes = [C.c_void_p, C.c_int]
    L.freedv_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.freedv_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.freedv_check_crc16_unpacked.restype = C.c_int
    L.freedv_check_crc16_unpacked.argtypes = [C.c_void_p, C.c_int]
    L.ofdm_generate_payload_data_bits.argtypes = [C.c_void_p, C.c_int]
    L.freedv_fsk_ldpc_rx_llrs.restype = C.c_int
    L.freedv_fsk_ldpc_rx_llrs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.freedv_fsk_ldpc_rx_info.argtypes = [C.c_void_p, C.c_void_p]
    L._ldpc_ready = True
    return L


def ldpc_code(name):
    """ldpc_codes_setup(): returns (struct LDPC, dict view of the tables for the tests)."""
    L = _ldpc_protos()
    if L.ldpc_codes_find(name.encode()) < 0:
        raise PiripError("unknown LDPC code " + name)
    l = LDPC()
    L.ldpc_codes_setup(C.byref(l), name.encode())
    nr = l.NumberParityBits * l.max_row_weight
    nc = l.NumberRowsHcols * l.max_col_weight
    view = dict(CodeLength=l.CodeLength, NumberParityBits=l.NumberParityBits, NumberRowsHcols=l.NumberRowsHcols,
                max_row_weight=l.max_row_weight, max_col_weight=l.max_col_weight,
                H_rows=np.ctypeslib.as_array(l.H_rows, (nr,)).copy(), H_cols=np.ctypeslib.as_array(l.H_cols, (nc,)).copy())
    return l, view


def ldpc_encode(l, ibits):
    L = _ldpc_protos()
    ibits = np.ascontiguousarray(ibits, np.uint8)
    p = np.zeros(l.NumberParityBits, np.uint8)
    L.ldpc_encode_frame(C.byref(l), _ptr(ibits), _ptr(p))
    return p


class LdpcPlan:
    def __init__(self, l, device=0):
        self.L = _ldpc_protos()
        self.N = l.CodeLength
        self._h = self.L.pirip_ldpc_plan_create(C.byref(l), device)
        if not self._h:
            raise _err("pirip_ldpc_plan_create")

    MODES = {"sum_product": 0, "sum_product_fast": 1, "min_sum": 2}

    def set_mode(self, mode):
        """'sum_product' (exact, default), 'sum_product_fast' or 'min_sum' (include/pirip_ldpc.h)."""
        self.L.pirip_ldpc_plan_set_mode.restype = C.c_int
        self.L.pirip_ldpc_plan_set_mode.argtypes = [C.c_void_p, C.c_int]
        rc = self.L.pirip_ldpc_plan_set_mode(self._h, self.MODES[mode])
        if rc != 0:
            raise _err("pirip_ldpc_plan_set_mode", rc)

    def decode(self, llr):
        llr = np.ascontiguousarray(llr, np.float32).reshape(-1, self.N)
        n = llr.shape[0]
        bits = np.zeros((n, self.N), np.uint8)
        it = np.zeros(n, np.int32)
        pcc = np.zeros(n, np.int32)
        rc = self.L.pirip_ldpc_decode_batch(self._h, _ptr(llr), n, _ptr(bits), _ptr(it), _ptr(pcc))
        if rc != 0:
            raise _err("pirip_ldpc_decode_batch", rc)
        return bits, it, pcc

    def close(self):
        if getattr(self, "_h", None):
            self.L.pirip_ldpc_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RxPlan:
    """pirip_fsk_ldpc_rx_plan_*: the coded receive chain of a batch of captures (config 4): demodulation with LLRs, UW sync,
    LDPC, CRC-16 on the device; per modem frame a status byte and the packed payload (the `rtl_fsk --code ... -b` records)."""

    def __init__(self, codename, Fs, Rs, M, P=8, Nsym=50, f1_tx=1200, tone_spacing=400, limits=None, mask=False, fmt=FMT_U8,
                 device=0, demod_mode=None, ldpc_mode=None):
        L = _ldpc_protos()
        lib()
        L.pirip_fsk_ldpc_rx_plan_create.restype = C.c_void_p
        L.pirip_fsk_ldpc_rx_plan_create.argtypes = [C.POINTER(FskCfg), C.c_char_p, C.c_int]
        L.pirip_fsk_ldpc_rx_plan_destroy.argtypes = [C.c_void_p]
        L.pirip_fsk_ldpc_rx_plan_fsk.restype = C.c_void_p
        L.pirip_fsk_ldpc_rx_plan_fsk.argtypes = [C.c_void_p]
        L.pirip_fsk_ldpc_rx_plan_ldpc.restype = C.c_void_p
        L.pirip_fsk_ldpc_rx_plan_ldpc.argtypes = [C.c_void_p]
        L.pirip_fsk_ldpc_rx_plan_payload_bytes.restype = C.c_int
        L.pirip_fsk_ldpc_rx_plan_payload_bytes.argtypes = [C.c_void_p]
        L.pirip_fsk_ldpc_rx_batch.restype = C.c_int
        L.pirip_fsk_ldpc_rx_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_size_t]
        L.pirip_fsk_ldpc_rx_batch_dev.restype = C.c_int
        L.pirip_fsk_ldpc_rx_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.pirip_ldpc_plan_set_mode.restype = C.c_int
        L.pirip_ldpc_plan_set_mode.argtypes = [C.c_void_p, C.c_int]
        self.L = L
        self.cfg = FskCfg(Fs=Fs, Rs=Rs, M=M, P=P, Nsym=Nsym, f1_tx=f1_tx, tone_spacing=tone_spacing,
                          est_min=limits[0] if limits else 0, est_max=limits[1] if limits else 0,
                          have_limits=1 if limits else 0, freq_est_type=1 if mask else 0, fmt=fmt, soft=1)
        self._h = L.pirip_fsk_ldpc_rx_plan_create(C.byref(self.cfg), codename.encode(), device)
        if not self._h:
            raise _err("pirip_fsk_ldpc_rx_plan_create")
        self.payload_bytes = L.pirip_fsk_ldpc_rx_plan_payload_bytes(self._h)
        self._fsk = L.pirip_fsk_ldpc_rx_plan_fsk(self._h)
        self._ldpc = L.pirip_fsk_ldpc_rx_plan_ldpc(self._h)
        self.nbits = lib().pirip_fsk_plan_nbits(self._fsk)
        self.bytes_per_sample = {FMT_U8: 2, FMT_S16: 4, FMT_F32: 8}[fmt]
        if demod_mode:
            rc = lib().pirip_fsk_plan_set_mode(self._fsk, {"exact": MODE_EXACT, "fast": MODE_FAST}[demod_mode])
            if rc != 0:
                raise _err("pirip_fsk_plan_set_mode", rc)
        if ldpc_mode:
            rc = L.pirip_ldpc_plan_set_mode(self._ldpc, LdpcPlan.MODES[ldpc_mode])
            if rc != 0:
                raise _err("pirip_ldpc_plan_set_mode", rc)

    def max_frames(self, samples_per_capture):
        return int(lib().pirip_fsk_plan_max_frames(self._fsk, samples_per_capture))

    def rx_batch(self, iq, want_info=False):
        """iq [n_captures, ...] one capture per row.  Returns dict(status[n, mf], payload[n, mf, bytes], nframes[n], info)."""
        iq = np.ascontiguousarray(iq)
        n = iq.shape[0]
        spc = iq.shape[1] * iq.dtype.itemsize // self.bytes_per_sample
        mf = self.max_frames(spc)
        status = np.zeros((n, mf), np.uint8)
        payload = np.zeros((n, mf, self.payload_bytes), np.uint8)
        nframes = np.zeros(n, np.int32)
        info = np.zeros((n, mf, 4), np.int32) if want_info else None
        rc = self.L.pirip_fsk_ldpc_rx_batch(self._h, _ptr(iq), n, spc, _ptr(status), _ptr(payload), _ptr(nframes), _ptr(info), mf)
        if rc != 0:
            raise _err("pirip_fsk_ldpc_rx_batch", rc)
        return dict(status=status, payload=payload, nframes=nframes, info=info, max_frames=mf)

    def rx_batch_dev(self, d_iq, n, spc, stride, d_llr, d_nframes, d_status, d_payload, max_frames, d_info=None, stream=None):
        rc = self.L.pirip_fsk_ldpc_rx_batch_dev(self._h, d_iq, n, spc, stride, d_llr, d_nframes, d_status, d_payload, d_info, max_frames, stream)
        if rc != 0:
            raise _err("pirip_fsk_ldpc_rx_batch_dev", rc)

    def close(self):
        if getattr(self, "_h", None):
            self.L.pirip_fsk_ldpc_rx_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FskLdpcModem:
    """freedv_open_advanced(FREEDV_MODE_FSK_LDPC) handle: framer ([R tx/rpitx_fsk.cpp:395]) and the
    receive-side deframer fed with LLRs."""

    def __init__(self, codename, M=2, Rs=10000, Fs=80000):
        self.L = _ldpc_protos()
        self._name = codename.encode()
        adv = FreedvAdvanced(0, M, Rs, Fs, 0, 0, self._name)
        self._h = self.L.freedv_open_advanced(FREEDV_MODE_FSK_LDPC, C.byref(adv))
        if not self._h:
            raise PiripError("freedv_open_advanced failed for code " + codename)
        self.data_bits = self.L.freedv_get_bits_per_modem_frame(self._h)
        self.bits_per_frame = self.L.freedv_tx_fsk_ldpc_bits_per_frame(self._h)

    def frame(self, payload_bits):
        payload_bits = np.ascontiguousarray(payload_bits, np.uint8)
        assert payload_bits.size == self.data_bits
        out = np.zeros(self.bits_per_frame, np.uint8)
        self.L.freedv_tx_fsk_ldpc_framer(self._h, _ptr(out), _ptr(payload_bits))
        return out

    def rx(self, llr):
        llr = np.ascontiguousarray(llr, np.float32)
        out = np.zeros(self.data_bits, np.uint8)
        st = self.L.freedv_fsk_ldpc_rx_llrs(self._h, _ptr(llr), llr.size, _ptr(out))
        return st, out

    def info(self):
        a = np.zeros(6, np.int32)
        self.L.freedv_fsk_ldpc_rx_info(self._h, _ptr(a))
        return dict(state=int(a[0]), uw_loc=int(a[1]), uw_err=int(a[2]), iter=int(a[3]), pcc=int(a[4]), nbits=int(a[5]))

    def close(self):
        if getattr(self, "_h", None):
            self.L.freedv_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
