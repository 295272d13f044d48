"""Debug aid: batch kernel (fsk_kernel2.cu) against the per-capture kernel (fsk_kernel.cu) on the same input."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pirip_b200 as pb
import synth

ncap = int(sys.argv[1]) if len(sys.argv) > 1 else 6
nbits = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
kw = dict(Fs=240000, Rs=10000, M=2, P=24, limits=(500, 25000))
caps = [synth.capture_u8(100 + i, nbits, 2, 240000, 10000, 10000, 10000, ebno_db=[None, 12.0, 9.0, 6.0][i % 4])[0] for i in range(ncap)]
n = min(len(c) for c in caps)
iq = np.stack([c[:n] for c in caps])
os.environ["PIRIP_FSK_K2"] = "0"
p0 = pb.FskPlan(**kw, device=0)
ref = p0.demod_batch(iq, want_stats=True, want_rx_filt=True)
p0.close()
os.environ["PIRIP_FSK_K2"] = "1"
p1 = pb.FskPlan(**kw, device=0)
out = p1.demod_batch(iq, want_stats=True, want_rx_filt=True)
p1.close()
print("nframes", ref["nframes"], out["nframes"])
for c in range(ncap):
    nf = int(ref["nframes"][c])
    nb = int((ref["bits"][c] != out["bits"][c]).sum())
    line = "cap %d: bit mismatches %d;" % (c, nb)
    for name in ref["stats"].dtype.names:
        a = np.ascontiguousarray(ref["stats"][c, :nf][name]).view(np.uint32).reshape(nf, -1)
        b = np.ascontiguousarray(out["stats"][c, :nf][name]).view(np.uint32).reshape(nf, -1)
        bad = np.nonzero((a != b).any(axis=1))[0]
        if bad.size:
            line += " %s:%d/%d(first %d)" % (name, bad.size, nf, bad[0])
    fa = ref["rx_filt"][c].view(np.uint32).reshape(-1, 100)[:nf]
    fb = out["rx_filt"][c].view(np.uint32).reshape(-1, 100)[:nf]
    badf = np.nonzero((fa != fb).any(axis=1))[0]
    line += " rx_filt frames bad %d/%d" % (badf.size, nf)
    print(line)
c = 0
nf = int(ref["nframes"][c])
print("norm_rx_timing old:", ref["stats"][c, :min(nf, 6)]["norm_rx_timing"])
print("norm_rx_timing new:", out["stats"][c, :min(nf, 6)]["norm_rx_timing"])
