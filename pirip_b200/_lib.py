"""ctypes loader for libpirip_b200.so and Python mirrors of the C-ABI structs."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FMT_U8, FMT_S16, FMT_F32 = 0, 1, 2


class PiripError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libpirip_b200.so")


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
    L.pirip_kernel_launches.restype = C.c_uint64
    L.pirip_last_error.restype = C.c_char_p
    L.pirip_version.restype = C.c_char_p
    _lib = L
    return L


def version():
    return lib().pirip_version().decode()


def kernel_launches():
    return int(lib().pirip_kernel_launches())


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
