#!/usr/bin/env python
"""Generate tests/golden/*.npz — small regression vectors for the IQ->bits path.

PARITY UNPINNED: the reference (/root/reference) holds no golden vector for this path and its
arithmetic lives in un-vendored codec2/csdr, so these vectors are produced by
the repo's own CPU oracle (oracle/, gcc -O2) from seeded synthetic input.  They pin the oracle
and the CUDA path against silent drift; they are not outputs of codec2.
Run from the repo root:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import synth  # noqa: E402

CASES = {
    # name: (oracle cfg kwargs, f1, spacing, nbits, ebnos)
    "c1_2fsk_p24": (dict(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000)), 10000, 10000, 800, [None, 7.0]),
    "c3_4fsk_soft": (dict(Fs=240000, Rs=10000, M=4, P=8, limits=(500, 45000), soft=True), 10000, 10000, 1200, [9.0, 5.0]),
    "mask_4fsk": (dict(Fs=200000, Rs=10000, M=4, P=20, tone_spacing=10000, mask=True, soft=True), 10000, 10000, 1200, [None, 8.0]),
}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (kw, f1, sp, nbits, ebnos) in CASES.items():
        cfg = orc.demod_cfg(kw["Fs"], kw["Rs"], kw["M"], P=kw["P"], limits=kw.get("limits"), mask=kw.get("mask", False),
                            tone_spacing=kw.get("tone_spacing", 400), soft=kw.get("soft", False))
        caps = [synth.capture_u8(7000 + i, nbits, kw["M"], kw["Fs"], kw["Rs"], f1, sp, ebno_db=e)[0]
                for i, e in enumerate(ebnos)]
        n = min(len(c) for c in caps)
        iq = np.stack([c[:n] for c in caps])
        res = [orc.fsk_demod_stream(cfg, iq[i]) for i in range(iq.shape[0])]
        arrays = dict(iq=iq, nframes=np.array([r["nframes"] for r in res], np.int32))
        for i, r in enumerate(res):
            arrays["bits_%d" % i] = r["bits"]
            arrays["stats_%d" % i] = r["stats"].view(np.uint8).reshape(r["nframes"], -1)
            arrays["rx_filt_%d" % i] = r["rx_filt"]
            if r["llr"] is not None:
                arrays["llr_%d" % i] = r["llr"]
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        print(name, iq.shape, [r["nframes"] for r in res])


if __name__ == "__main__":
    main()
