"""Independent float64 model of the FSK demodulator, written from SURVEY.md Appendix A.1 - A.3 (the survey's description of
codec2's fsk_create_core / fsk_demod_freq_est / fsk_demod_core), NOT from oracle/orc_fsk.c: a second opinion on the oracle's
transcription.  numpy, double precision, vectorised where the description allows; the tests compare the
oracle with it within float32 rounding.  Test infrastructure."""
import math

import numpy as np


class Model:
    def __init__(self, Fs, Rs, M, P=8, Nsym=50, tone_spacing=400, limits=None, mask=False):
        assert Fs % Rs == 0 and (Fs // Rs) % P == 0 and P >= 4 and M in (2, 4)
        self.Fs, self.Rs, self.M, self.P, self.Nsym = Fs, Rs, M, P, Nsym
        self.Ts = Fs // Rs
        self.N = self.Ts * Nsym
        self.Nmem = self.N + 2 * self.Ts
        self.nin = self.N
        self.Nbits = Nsym * (1 if M == 2 else 2)
        self.Ndft = 2 ** int(math.ceil(math.log2(Fs / (0.1 * Rs))))
        self.tc = 0.1
        self.est_min, self.est_max = limits if limits else (0, Fs)  # the oracle's documented default window
        self.est_space = int(0.75 * Rs)
        self.tone_spacing = tone_spacing
        self.mask = mask
        self.phi = np.ones(M, complex)
        self.f_dc = np.zeros((M, self.Nmem), complex)
        self.Sf = np.zeros(self.Ndft)
        self.norm_rx_timing = 0.0
        self.ppm = 0.0
        i = np.arange(self.Ndft)
        self.hann = 0.5 - 0.5 * np.cos(2 * np.pi * i / (self.Ndft - 1))

    def freq_est(self, x):
        Ndft, Fs, M = self.Ndft, self.Fs, self.M
        st = max(int(self.est_min * Ndft / Fs) + Ndft // 2, 0)  # C integer division truncates towards zero
        en = min(int(self.est_max * Ndft / Fs) + Ndft // 2, Ndft)
        f_zero = int(self.est_space * Ndft / Fs)
        numffts = len(x) // (Ndft // 2) - 1
        for j in range(numffts):
            a = j * Ndft // 2
            X = np.fft.fft(self.hann * x[a:a + Ndft])
            X = np.concatenate([X[Ndft // 2:], X[:Ndft // 2]])
            self.Sf = (1 - self.tc) * self.Sf + self.tc * np.abs(X)
        tmp = self.Sf.copy()
        freqi = []
        for m in range(M):
            seg = tmp[st:en]
            imax = st + int(np.argmax(seg)) if seg.max() > 0 else 0
            lo, hi = max(imax - f_zero, 0), min(imax + f_zero, Ndft)
            tmp[lo:hi] = 0
            freqi.append(imax - Ndft // 2)
        freqi.sort()
        f_est = [fi * Fs / Ndft for fi in freqi]
        mask = np.zeros(Ndft)
        mask[0:3] = 1
        b = 0
        for m in range(1, M):
            b = int(round(m * self.tone_spacing * Ndft / Fs)) - 1
            mask[b:b + 3] = 1
        len_mask = b + 3
        b_max, corr_max = st, 0.0
        for bb in range(st, en - len_mask):
            c = float(np.dot(mask[:len_mask], self.Sf[bb:bb + len_mask]))
            if c > corr_max:
                corr_max, b_max = c, bb
        foff = int((b_max - Ndft // 2) * Fs / Ndft)  # a C int expression upstream: truncated
        f2_est = [foff + m * self.tone_spacing for m in range(M)]
        return f_est, f2_est, freqi

    def demod(self, x):
        """x: complex128[nin].  Returns dict(bits, rx_filt [M, Nsym], margin [Nsym], f_est, norm_rx_timing, nin_next)."""
        M, Ts, P, Nsym, N, Nmem = self.M, self.Ts, self.P, self.Nsym, self.N, self.Nmem
        nin = len(x)
        assert nin == self.nin
        f_est, f2_est, freqi = self.freq_est(x)
        f = f2_est if self.mask else f_est
        nold = Nmem - nin
        self.f_dc[:, :nold] = self.f_dc[:, nin:].copy()
        j = np.arange(1, nin + 1)
        for m in range(M):
            d = np.exp(2j * np.pi * f[m] / self.Fs)
            ph = self.phi[m] * d ** j  # phi_c *= d, then used: the oscillator of sample j is phi * d^(j+1)
            self.f_dc[m, nold:] = x * np.conj(ph)
            self.phi[m] = ph[-1] / abs(ph[-1])
        nint = (Nsym + 1) * P
        cs = np.concatenate([np.zeros((M, 1), complex), np.cumsum(self.f_dc, axis=1)], axis=1)
        s = np.arange(nint) * Ts // P
        f_int = cs[:, s + Ts] - cs[:, s]
        e = (np.abs(f_int) ** 2).sum(axis=0)
        t_c = np.sum(e * np.exp(2j * np.pi * np.arange(nint) / P))
        norm = math.atan2(t_c.imag, t_c.real) / (2 * np.pi)
        rx_timing = norm * P
        dn = norm - self.norm_rx_timing
        self.norm_rx_timing = norm
        if abs(dn) < 0.2:
            self.ppm = 0.9 * self.ppm + 0.1 * (1e6 * dn / Nsym)
        self.nin = N + Ts // 4 if norm > 0.25 else (N - Ts // 4 if norm < -0.25 else N)
        low, high = math.floor(rx_timing), math.ceil(rx_timing)
        fract = rx_timing - low
        st = (np.arange(Nsym) + 1) * P
        t = (1 - fract) * f_int[:, st + low] + fract * f_int[:, st + high]
        en = np.abs(t) ** 2
        sym = np.argmax(en, axis=0)
        srt = np.sort(en, axis=0)
        margin = (srt[-1] - srt[-2]) / np.maximum(srt[-1] + srt[-2], 1e-300)
        if M == 2:
            bits = (sym == 1).astype(np.uint8)
        else:
            bits = np.stack([(sym >> 1) & 1, sym & 1], axis=1).ravel().astype(np.uint8)
        return dict(bits=bits, rx_filt=np.sqrt(en), margin=margin, f_est=f_est, f2_est=f2_est, norm_rx_timing=norm, nin_next=self.nin,
                    timing_margin=abs(abs(norm) - 0.25))


def demod_stream(model, x):
    """The fsk_demod main loop (A.4) over complex samples x; stops on a short read."""
    pos, out = 0, []
    while pos + model.nin <= len(x):
        n = model.nin
        out.append(model.demod(x[pos:pos + n]))
        pos += n
    return out
