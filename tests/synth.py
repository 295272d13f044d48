"""Synthetic IQ generation for tests and bench (SURVEY.md 8d "Synthetic inputs").

Host-side, numpy, float64 phase accumulator: continuous-phase M-FSK with tone index =
symbol value and ascending frequency ([R tx/rpitx_fsk.cpp:127-143]), complex AWGN from a
counter-based Philox stream, 8-bit quantisation to the interleaved u8 I,Q format rtl_sdr
emits ([R test/loopback_rtl_sdr.sh:16]).  This is input generation, not part of the
hot path, and is shared by the oracle and the CUDA path (both get the same bytes).
"""
import numpy as np

BASE_SEED = 0x50495250  # "PIRP"

# fsk_get_test_bits frame: srand(158324), 100 x rand()&1 under glibc;
# tests/test_oracle_framing.py checks the oracle's rand() frame equals this constant.
TEST_FRAME_100 = np.array(list(map(int,
    "0101111110001000101010010010010100011000011111001110111111100111110011100011110001110001111000110101")),
    np.uint8)


def rng(seed, stream=0):
    return np.random.Generator(np.random.Philox(key=[(BASE_SEED + seed) & 0xFFFFFFFFFFFFFFFF, stream]))


def test_bits(nbits, frame=TEST_FRAME_100):
    """fsk_get_test_bits: the 100-bit frame repeated to nbits ([R README.md:101])."""
    reps = -(-nbits // frame.size)
    return np.tile(frame, reps)[:nbits].copy()


test_bits.__test__ = False


def bits_to_syms(bits, M):
    bits = np.asarray(bits, np.int64)
    if M == 2:
        return bits
    assert bits.size % 2 == 0
    return bits[0::2] * 2 + bits[1::2]  # MSB first ([R tx/rpitx_fsk.cpp:133-138])


def fsk_mod(bits, M, Fs, Rs, f1, spacing, *, nsamples=None, timing_offset=0.0, freq_offset=0.0, ppm=0.0,
            phase0=0.0, lead_in=0):
    """Complex baseband CPFSK, unit amplitude.  timing_offset: fractional-sample delay of the
    symbol clock; ppm: sample-clock error (receiver sees Ts*(1+ppm*1e-6) samples/symbol);
    lead_in: samples of tone 0 before the first symbol."""
    syms = bits_to_syms(bits, M)
    Ts = Fs / Rs * (1.0 + ppm * 1e-6)
    if nsamples is None:
        nsamples = int(np.floor(syms.size * Ts + timing_offset)) + lead_in
    n = np.arange(nsamples, dtype=np.float64)
    k = np.floor((n - lead_in - timing_offset) / Ts).astype(np.int64)
    sym = np.where(k < 0, 0, syms[np.clip(k, 0, syms.size - 1)])
    sym = np.where(k >= syms.size, 0, sym)
    f = (f1 + freq_offset + spacing * sym) / Fs
    phase = phase0 + 2 * np.pi * np.cumsum(f)
    return np.exp(1j * phase)


def add_awgn(x, ebno_db, Ts, M, g):
    """Eb/N0 for a unit-amplitude complex signal: Es = Ts, N0 = sigma^2 (complex variance)."""
    if ebno_db is None or np.isinf(ebno_db):
        return x
    esno = 10 ** (ebno_db / 10) * np.log2(M)
    sigma2 = Ts / esno
    n = g.standard_normal(x.size) + 1j * g.standard_normal(x.size)
    return x + n * np.sqrt(sigma2 / 2)


def to_u8(x, amp=100.0, noise_floor=0.0, g=None):
    """Scale to +/-amp LSB about mid-scale, round to nearest, clip: rtl_sdr's u8 I,Q bytes."""
    y = x * amp
    if noise_floor > 0 and g is not None:
        y = y + noise_floor * (g.standard_normal(x.size) + 1j * g.standard_normal(x.size))
    out = np.empty(2 * x.size, np.uint8)
    out[0::2] = np.clip(np.rint(127.5 + y.real), 0, 255).astype(np.uint8)
    out[1::2] = np.clip(np.rint(127.5 + y.imag), 0, 255).astype(np.uint8)
    return out


def to_s16(x, amp=8000.0):
    out = np.empty(2 * x.size, np.int16)
    out[0::2] = np.clip(np.rint(x.real * amp), -32768, 32767).astype(np.int16)
    out[1::2] = np.clip(np.rint(x.imag * amp), -32768, 32767).astype(np.int16)
    return out


def capture_u8(seed, nbits, M, Fs, Rs, f1, spacing, ebno_db=None, impair=True, amp=None, bits=None):
    """One capture of the config-5 kind: test bits, random phase / timing / frequency / clock
    offsets, AWGN, u8."""
    g = rng(seed)
    if bits is None:
        bits = test_bits(nbits)
    Ts = Fs // Rs
    if impair:
        phase0 = g.uniform(0, 2 * np.pi)
        toff = g.uniform(0, Ts)
        foff = g.uniform(-0.1, 0.1) * Rs
        ppm = g.uniform(-50, 50)
    else:
        phase0 = toff = foff = ppm = 0.0
    x = fsk_mod(bits, M, Fs, Rs, f1, spacing, timing_offset=toff, freq_offset=foff, ppm=ppm, phase0=phase0)
    if amp is None:
        # leave head-room for the noise so clipping is rare
        amp = 100.0 if ebno_db is None else min(100.0, 110.0 / (1.0 + 3.0 * np.sqrt(Ts / (2 * 10 ** (ebno_db / 10) * np.log2(M)))))
    x = add_awgn(x, ebno_db, Ts, M, g)
    return to_u8(x, amp=amp), bits


def ber_vs_stream(rx_bits, frame=TEST_FRAME_100):
    """fsk_put_test_bits-style error count: align the known repeating frame to the received
    stream (best of all cyclic phases) and count bit errors after a settling prefix."""
    rx = np.asarray(rx_bits, np.uint8)
    L = frame.size
    skip = min(rx.size // 4, 4 * L)
    body = rx[skip:]
    n = (body.size // L) * L
    if n == 0:
        return 1.0, 0
    body = body[:n].reshape(-1, L)
    best = None
    for ph in range(L):
        ref = np.roll(frame, -ph)
        e = int((body != ref).sum())
        if best is None or e < best:
            best = e
    return best / n, n
